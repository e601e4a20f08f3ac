"""ctypes front-end of the CPU ORACLE (test infrastructure — never imported by the product).

Builds the literal ``OraProto`` table from a protocol dict and runs oracle/_build/libsd_oracle.so
on packed batches.  Results come back as canonical tuples
``(status, [(protocol_id, payload, bit_length), ...])`` per message, the same form
``oracle/ref_import.py`` produces from the real reference.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path
from typing import Any, Dict, List, Sequence, Tuple

import numpy as np

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "_build" / "libsd_oracle.so"
ORA_MAXLIST = 16

PD_IDS = {
    "postDemo_EM": 1, "postDemo_Revolt": 2, "postDemo_FS20": 3, "postDemo_FHT80": 4,
    "postDemo_FHT80TF": 5, "postDemo_WS2000": 6, "postDemo_WS7035": 7, "postDemo_WS7053": 8,
    "postDemo_lengtnPrefix": 9,
}
# method name (after the last '.') -> ORA_M_*; note 'mcraw' (helpers.py:90) vs 'mcRaw' (manchester.py:588)
METHOD_IDS = {
    "mcBit2Funkbus": 1, "mcBit2Sainlogic": 2, "mcBit2AS": 3, "mcBit2Hideki": 4, "mcBit2Maverick": 5,
    "mcBit2OSV1": 6, "mcBit2OSV2o3": 7, "mcBit2OSPIR": 8, "mcRaw": 9, "mcraw": 10, "mcBit2TFA": 11,
    "mcBit2Grothe": 12, "mcBit2SomfyRTS": 13,
    "ConvBresser_lightning": 14, "ConvBresser_5in1": 15, "ConvBresser_6in1": 16, "ConvBresser_7in1": 17,
    "ConvPCA301": 18, "ConvKoppFreeControl": 19, "ConvLaCrosse": 20,
}
M_UNKNOWN = 21
STATUS_NAMES = {0: "ok", 1: "IndexError", 2: "TypeError", 3: "ValueError"}


class OraProto(C.Structure):
    _fields_ = [
        ("id", C.c_char * 16),
        ("has_clockabs", C.c_int32), ("clockabs", C.c_double),
        ("sync_kind", C.c_int32),
        ("nsync", C.c_int32), ("sync", C.c_double * ORA_MAXLIST),
        ("start_is_list", C.c_int32),
        ("nstart", C.c_int32), ("start", C.c_double * ORA_MAXLIST),
        ("none", C.c_int32), ("one", C.c_double * ORA_MAXLIST),
        ("nzero", C.c_int32), ("zero", C.c_double * ORA_MAXLIST),
        ("nfloat", C.c_int32), ("flt", C.c_double * ORA_MAXLIST),
        ("has_length_min", C.c_int32), ("length_min", C.c_int32),
        ("has_length_max", C.c_int32), ("length_max", C.c_int32),
        ("length_max_truthy", C.c_int32),
        ("paddingbits", C.c_int32),
        ("postdemod", C.c_int32),
        ("reconstruct", C.c_int32),
        ("dispatch_bin", C.c_int32),
        ("remove_zero", C.c_int32),
        ("active", C.c_int32),
        ("has_modulematch", C.c_int32),
        ("modulematch", C.c_char * 96),
        ("preamble", C.c_char * 32),
        ("postamble", C.c_char * 16),
        ("method", C.c_int32),
        ("has_clockrange", C.c_int32), ("clock_min", C.c_int32), ("clock_max", C.c_int32),
        ("polarity_invert", C.c_int32),
        ("length_max_is_str", C.c_int32),
    ]


class OraHit(C.Structure):
    _fields_ = [("msg", C.c_int32), ("proto", C.c_int32), ("bit_length", C.c_int32),
                ("payload_off", C.c_int32), ("payload_len", C.c_int32)]


ORAHIT_DTYPE = np.dtype([("msg", "<i4"), ("proto", "<i4"), ("bit_length", "<i4"),
                         ("payload_off", "<i4"), ("payload_len", "<i4")])


def build() -> Path:
    """Compile the oracle with gcc (make -C oracle)."""
    subprocess.run(["make", "-s", "-C", str(HERE)], check=True)
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            build()
        L = C.CDLL(str(LIB_PATH))
        L.ora_round1.restype = C.c_double
        L.ora_round1.argtypes = [C.c_double]
        L.ora_round1_printf.restype = C.c_double
        L.ora_round1_printf.argtypes = [C.c_double]
        L.ora_tolerance.restype = C.c_double
        L.ora_tolerance.argtypes = [C.c_double]
        L.ora_demod_pulse.restype = C.c_int
        L.ora_demod_hex.restype = C.c_int
        _lib = L
    return _lib


def _set_list(p: OraProto, name: str, values, count_field: str) -> None:
    if not values:
        setattr(p, count_field, 0)
        return
    if len(values) > ORA_MAXLIST:
        raise NotImplementedError(f"{name} longer than {ORA_MAXLIST}")
    arr = getattr(p, name)
    for i, v in enumerate(values):
        arr[i] = float(v)
    setattr(p, count_field, len(values))


def make_table(protocols: Dict[str, Dict[str, Any]]):
    """Literal transcription of the protocol dict into OraProto[] (table order)."""
    n = len(protocols)
    tab = (OraProto * n)()
    for i, (pid, pr) in enumerate(protocols.items()):
        p = tab[i]
        p.id = pid.encode()
        if "clockabs" in pr:
            p.has_clockabs = 1
            p.clockabs = float(pr["clockabs"])
        if "sync" in pr:
            sv = pr["sync"]
            if not sv:
                p.sync_kind = 3
            else:
                try:
                    vals = [float(x) for x in sv]
                    p.sync_kind = 1
                    _set_list(p, "sync", vals, "nsync")
                except (ValueError, TypeError):
                    p.sync_kind = 2
        st = pr.get("start")
        if st and isinstance(st, list):
            p.start_is_list = 1
            _set_list(p, "start", st, "nstart")
        _set_list(p, "one", pr.get("one"), "none")
        _set_list(p, "zero", pr.get("zero"), "nzero")
        _set_list(p, "flt", pr.get("float"), "nfloat")
        if pr.get("length_min") is not None:
            p.has_length_min = 1
            p.length_min = int(pr["length_min"])
        if pr.get("length_max") is not None:
            p.has_length_max = 1
            p.length_max = int(pr["length_max"])
            p.length_max_truthy = 1 if pr["length_max"] else 0
            p.length_max_is_str = 1 if isinstance(pr["length_max"], str) else 0
        p.paddingbits = int(pr.get("paddingbits", 4))
        pd = pr.get("postDemodulation")
        if pd:
            p.postdemod = PD_IDS.get(pd.split(".")[-1], 0)
        p.reconstruct = 1 if pr.get("reconstructBit") else 0
        p.dispatch_bin = 1 if int(pr.get("dispatchBin", 0)) == 1 else 0
        p.remove_zero = 1 if pr.get("remove_zero", 0) else 0
        p.active = 1 if pr.get("active", True) else 0
        mm = pr.get("modulematch")
        if mm:
            p.has_modulematch = 1
            p.modulematch = mm.encode()
        p.preamble = str(pr.get("preamble", "")).encode()
        p.postamble = str(pr.get("postamble", "")).encode()
        meth = pr.get("method")
        if meth:
            p.method = METHOD_IDS.get(meth.split(".")[-1], M_UNKNOWN)
        cr = pr.get("clockrange")
        if cr and len(cr) >= 2:
            p.has_clockrange = 1
            p.clock_min, p.clock_max = int(cr[0]), int(cr[1])
        p.polarity_invert = 1 if pr.get("polarity", "") == "invert" else 0
    return tab


Canonical = Tuple[str, List[Tuple[str, str, int]]]


def _collect(n: int, status: np.ndarray, hits: np.ndarray, pool: bytes, ids: List[str]) -> List[Canonical]:
    out: List[Canonical] = [(STATUS_NAMES[int(s)], []) for s in status]
    for h in hits:
        o, l = int(h["payload_off"]), int(h["payload_len"])
        out[int(h["msg"])][1].append((ids[int(h["proto"])], pool[o : o + l].decode("latin-1"), int(h["bit_length"])))
    return out


class Oracle:
    """CPU oracle bound to one protocol table."""

    def __init__(self, protocols: Dict[str, Dict[str, Any]]):
        self.ids = list(protocols.keys())
        self.tab = make_table(protocols)
        self.n = len(self.ids)

    def run_pulse_raw(self, batch, nthreads: int = 1):
        """Run MS/MU; returns (status u8[n], hits structured array, pool bytes)."""
        L = lib()
        n = batch.n
        status = np.zeros(n, dtype=np.uint8)
        hits_cap, pool_cap = max(1024, 4 * n), max(1 << 16, 128 * n)
        msgs = np.ascontiguousarray(batch.msgs)
        digits = np.ascontiguousarray(batch.digits)
        while True:
            hits = np.zeros(hits_cap, dtype=ORAHIT_DTYPE)
            pool = np.zeros(pool_cap, dtype=np.uint8)
            nh, pu = C.c_int64(0), C.c_int64(0)
            rc = L.ora_demod_pulse(self.tab, C.c_int(self.n), C.c_int(batch.kind),
                                   C.c_void_p(msgs.ctypes.data), C.c_void_p(digits.ctypes.data), C.c_int64(n),
                                   C.c_void_p(status.ctypes.data), C.c_void_p(hits.ctypes.data), C.c_int64(hits_cap),
                                   C.c_void_p(pool.ctypes.data), C.c_int64(pool_cap), C.byref(nh), C.byref(pu),
                                   C.c_int(nthreads))
            if rc == -3:
                hits_cap, pool_cap = max(hits_cap, nh.value + 16), max(pool_cap, pu.value + 16)
                continue
            if rc != 0:
                raise RuntimeError(f"oracle failed rc={rc}")
            return status, hits[: nh.value], pool[: pu.value].tobytes()

    def run_pulse(self, batch, nthreads: int = 1) -> List[Canonical]:
        status, hits, pool = self.run_pulse_raw(batch, nthreads)
        return _collect(batch.n, status, hits, pool, self.ids)

    def run_hex_raw(self, batch, mc_repaired: bool = True, nthreads: int = 1):
        L = lib()
        n = batch.n
        status = np.zeros(n, dtype=np.uint8)
        hits_cap, pool_cap = max(1024, 4 * n), max(1 << 16, 160 * n)
        msgs = np.ascontiguousarray(batch.msgs)
        digits = np.ascontiguousarray(batch.digits)
        while True:
            hits = np.zeros(hits_cap, dtype=ORAHIT_DTYPE)
            pool = np.zeros(pool_cap, dtype=np.uint8)
            nh, pu = C.c_int64(0), C.c_int64(0)
            rc = L.ora_demod_hex(self.tab, C.c_int(self.n), C.c_int(batch.kind), C.c_int(1 if mc_repaired else 0),
                                 C.c_void_p(msgs.ctypes.data), C.c_void_p(digits.ctypes.data), C.c_int64(n),
                                 C.c_void_p(status.ctypes.data), C.c_void_p(hits.ctypes.data), C.c_int64(hits_cap),
                                 C.c_void_p(pool.ctypes.data), C.c_int64(pool_cap), C.byref(nh), C.byref(pu),
                                 C.c_int(nthreads))
            if rc == -3:
                hits_cap, pool_cap = max(hits_cap, nh.value + 16), max(pool_cap, pu.value + 16)
                continue
            if rc != 0:
                raise RuntimeError(f"oracle failed rc={rc}")
            return status, hits[: nh.value], pool[: pu.value].tobytes()

    def run_hex(self, batch, mc_repaired: bool = True, nthreads: int = 1) -> List[Canonical]:
        status, hits, pool = self.run_hex_raw(batch, mc_repaired, nthreads)
        return _collect(batch.n, status, hits, pool, self.ids)
