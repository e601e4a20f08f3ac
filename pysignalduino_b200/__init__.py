"""pysignalduino_b200 — B200-native batch demodulator behind PySignalduino's SDProtocols API.

``from pysignalduino_b200 import SDProtocols`` is the drop-in for
``from sd_protocols import SDProtocols`` (sd_protocols/__init__.py:2).
"""
from .sd_protocols import SDProtocols  # noqa: F401
from .pack import DomainError  # noqa: F401

VERSION = "1.0"

__all__ = ["SDProtocols", "DomainError", "VERSION"]
