"""pysignalduino_b200 — B200-native batch demodulator behind PySignalduino's SDProtocols API.

``from pysignalduino_b200 import SDProtocols`` is the drop-in for
``from sd_protocols import SDProtocols`` (sd_protocols/__init__.py:2).
"""
from .sd_protocols import SDProtocols  # noqa: F401
from .pack import DomainError  # noqa: F401
from .parser import DecodedMessage, RawFrame, SignalParser  # noqa: F401  (signalduino/parser/__init__.py:17, types.py:13-31)

VERSION = "1.0"

__all__ = ["SDProtocols", "SignalParser", "RawFrame", "DecodedMessage", "DomainError", "VERSION"]
