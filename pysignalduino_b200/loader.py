"""Method loader of the drop-in package: ``resolve_method`` / ``run_method`` / ``protocols``.

Mirrors ``sd_protocols/loader.py:15-72``: the ``method`` strings of the protocol table (``"manchester.mcBit2Grothe"``,
``"helpers.ConvBresser_6in1"`` …) resolve to bound methods of ONE module-level ``SDProtocols`` instance — whose
decoders run on the GPU (unit ops of libsdb200.so).  Importing this module does not touch the GPU; the first call of a
resolved decoder creates the engine (and fails loudly without a device: there is no CPU fallback).
"""
from __future__ import annotations

from .protocol_data import load_protocol_table
from .sd_protocols import SDProtocols

# sd_protocols/loader.py:6-10: the raw table, as loaded (no `active` / `name` defaults applied)
protocols = load_protocol_table()

# sd_protocols/loader.py:13. MC decoders run "repaired" here: the loader calls them directly with their own signature
# (method(name=..., bit_data=..., protocol_id=..., mcbitnum=...)), which works in the reference as shipped as well.
_protocol_handler = SDProtocols()


def resolve_method(path: str):
    """``'module.method'`` -> the bound ``SDProtocols`` method (sd_protocols/loader.py:15-48).

    Only the part after the first dot selects the method (every mixin's methods live on the one class); a path without
    a dot raises ``ValueError``, an unknown method name ``AttributeError``."""
    if "." not in path:
        raise ValueError(f"Invalid method path: {path}. Expected format: 'module.method'")
    _module_name, method_name = path.split(".", 1)
    method = getattr(_protocol_handler, method_name, None)
    if method is None:
        raise AttributeError(f"Method '{method_name}' not found in SDProtocols (path: {path})")
    return method


def run_method(pid, *args, **kwargs):
    """Call the table's ``method`` of protocol ``pid`` (sd_protocols/loader.py:50-72)."""
    proto = protocols.get(str(pid))
    if not proto or "method" not in proto:
        raise ValueError(f"Kein method-handler für Protokoll {pid}")
    return resolve_method(proto["method"])(*args, **kwargs)
