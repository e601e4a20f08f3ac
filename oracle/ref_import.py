"""Import the REAL reference: /root/reference in the build container, else the unmodified copy that
oracle/make_ref.py installed into the git-ignored oracle/_ref/ (which travels to the GPU box).

Used by oracle/validate_vs_reference.py and tests/golden/make_golden*.py to pin the C oracle and to generate the
committed golden fixtures, and by bench.py's cpu_baseline / --impl reference legs to TIME the reference's own Python
path on the GPU box's host cores.  Never imported by the product.
"""
from __future__ import annotations

import sys
from pathlib import Path
from typing import Any, Dict, List, Tuple

REFERENCE_ROOT = Path("/root/reference")
if not (REFERENCE_ROOT / "sd_protocols" / "sd_protocols.py").exists():
    REFERENCE_ROOT = Path(__file__).resolve().parent / "_ref"


def available() -> bool:
    return (REFERENCE_ROOT / "sd_protocols" / "sd_protocols.py").exists()


def reference_class():
    if not available():
        raise RuntimeError("reference not present: neither /root/reference nor oracle/_ref (python oracle/make_ref.py)")
    if str(REFERENCE_ROOT) not in sys.path:
        sys.path.insert(0, str(REFERENCE_ROOT))
    from sd_protocols import SDProtocols  # type: ignore

    return SDProtocols


def canonical(results: List[Dict[str, Any]]) -> List[Tuple[str, str, int]]:
    out = []
    for r in results:
        meta = r.get("meta") or {}
        out.append((str(r["protocol_id"]), str(r["payload"]), int(meta.get("bit_length", -1))))
    return out


def ref_demodulate(proto, msg: Dict[str, Any], msg_type: str):
    """-> (status, [(protocol_id, payload, bit_length)...]) with exceptions mapped to their type name."""
    try:
        return ("ok", canonical(proto.demodulate(dict(msg), msg_type)))
    except Exception as e:  # noqa: BLE001 - the exception type IS the result
        return (type(e).__name__, [])


_REPAIRED = None


def repaired_class():
    """The reference class with EXACTLY the two one-line MC repairs of SURVEY.md §8c applied in memory:

        manchester.py:83   clock_min, clock_max = clockrange, clockrange
                       ->  clock_min, clock_max = clockrange[0], clockrange[1]
        manchester.py:120  rcode, res = method_func(self, name, bit_data, protocol_id, len(bit_data))
                       ->  rcode, res = method_func(name, bit_data, protocol_id, len(bit_data))

    Nothing is written to disk; /root/reference stays untouched.
    """
    global _REPAIRED
    if _REPAIRED is not None:
        return _REPAIRED
    base = reference_class()
    import types

    import sd_protocols.manchester as man  # type: ignore

    src = (REFERENCE_ROOT / "sd_protocols" / "manchester.py").read_text(encoding="utf-8")
    a = "clock_min, clock_max = clockrange, clockrange"
    b = "rcode, res = method_func(self, name, bit_data, protocol_id, len(bit_data))"
    assert src.count(a) == 1 and src.count(b) == 1, "reference manchester.py changed: repairs do not apply"
    src = src.replace(a, "clock_min, clock_max = clockrange[0], clockrange[1]")
    src = src.replace(b, "rcode, res = method_func(name, bit_data, protocol_id, len(bit_data))")
    mod = types.ModuleType("sd_protocols.manchester_repaired")
    mod.__dict__["__name__"] = "sd_protocols.manchester_repaired"
    mod.__dict__["__package__"] = "sd_protocols"
    exec(compile(src, "manchester_repaired.py", "exec"), mod.__dict__)

    class SDProtocolsRepaired(base):  # type: ignore
        _demodulate_mc_data = mod.ManchesterMixin._demodulate_mc_data

    _REPAIRED = SDProtocolsRepaired
    return _REPAIRED


def reference_parser_module():
    """signalduino.parser of the reference WITHOUT running signalduino/__init__.py (it imports the asyncio controller
    and its third-party dependencies: aiomqtt, serial, ... — SURVEY App. D).  A bare package object with the right
    __path__ lets `signalduino.parser`, `.types` and `.exceptions` import normally."""
    import importlib
    import types

    reference_class()                                   # sd_protocols on sys.path
    if "signalduino" not in sys.modules:
        pkg = types.ModuleType("signalduino")
        pkg.__path__ = [str(REFERENCE_ROOT / "signalduino")]
        sys.modules["signalduino"] = pkg
    return importlib.import_module("signalduino.parser")
