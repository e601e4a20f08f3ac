"""Deterministic synthetic corpora for the BASELINE.json workloads (bench / test infrastructure).

Thin ctypes wrapper over corpus/_build/libsd_corpus.so; produces packed batches
(pysignalduino_b200.pack.PulseBatch / HexBatch) directly, so a 10 M-message corpus never
exists as Python dicts.  ``batch_to_dicts`` goes the other way for the CPU reference.

Seeds: MS 0x5D01, MU 0x5D02, MC/MN 0x5D03, mixed 0x5D05.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path
from typing import Any, Dict, List

import numpy as np

from pysignalduino_b200 import pack

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "_build" / "libsd_corpus.so"
GEN_MAXLIST = 16

SEED_MS, SEED_MU, SEED_MC, SEED_MIXED = 0x5D01, 0x5D02, 0x5D03, 0x5D05

METHOD_IDS = {
    "mcBit2Funkbus": 1, "mcBit2Sainlogic": 2, "mcBit2AS": 3, "mcBit2Hideki": 4, "mcBit2Maverick": 5,
    "mcBit2OSV1": 6, "mcBit2OSV2o3": 7, "mcBit2OSPIR": 8, "mcRaw": 9, "mcraw": 10, "mcBit2TFA": 11,
    "mcBit2Grothe": 12, "mcBit2SomfyRTS": 13,
    "ConvBresser_lightning": 14, "ConvBresser_5in1": 15, "ConvBresser_6in1": 16, "ConvBresser_7in1": 17,
    "ConvPCA301": 18, "ConvKoppFreeControl": 19, "ConvLaCrosse": 20,
}


class GenProto(C.Structure):
    _fields_ = [
        ("is_ms", C.c_int32), ("has_clockabs", C.c_int32), ("clockabs", C.c_double),
        ("nsync", C.c_int32), ("sync", C.c_double * GEN_MAXLIST),
        ("nstart", C.c_int32), ("start", C.c_double * GEN_MAXLIST),
        ("none", C.c_int32), ("one", C.c_double * GEN_MAXLIST),
        ("nzero", C.c_int32), ("zero", C.c_double * GEN_MAXLIST),
        ("nfloat", C.c_int32), ("flt", C.c_double * GEN_MAXLIST),
        ("npause", C.c_int32), ("pause", C.c_double * GEN_MAXLIST),
        ("nend", C.c_int32), ("end", C.c_double * GEN_MAXLIST),
        ("length_min", C.c_int32), ("length_max", C.c_int32), ("reconstruct", C.c_int32),
        ("method", C.c_int32), ("clock_min", C.c_int32), ("clock_max", C.c_int32),
        ("polarity_invert", C.c_int32), ("table_index", C.c_int32), ("is_119", C.c_int32),
    ]


def build() -> Path:
    subprocess.run(["make", "-s", "-C", str(HERE)], check=True)
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            build()
        _lib = C.CDLL(str(LIB_PATH))
        _lib.sdc_gen_pulse.restype = C.c_int
        _lib.sdc_free.argtypes = [C.c_void_p]
    return _lib


def _numlist(v):
    if isinstance(v, list) and v:
        try:
            return [float(x) for x in v]
        except (TypeError, ValueError):
            return []
    return []


class RssiCodes:
    """Lazy ``msg_data.get('R')`` view over generator codes (-1 absent, -2 the corrupt "1q")."""

    def __init__(self, codes: np.ndarray):
        self.codes = codes

    def __len__(self):
        return len(self.codes)

    def __getitem__(self, i):
        c = int(self.codes[i])
        return None if c == -1 else ("1q" if c == -2 else str(c))


class Corpus:
    """Generator bound to one protocol table."""

    def __init__(self, protocols: Dict[str, Dict[str, Any]]):
        self.ids = list(protocols)
        n = len(self.ids)
        self.tab = (GenProto * n)()
        ms_ids, mu_ids = [], []
        for i, (pid, pr) in enumerate(protocols.items()):
            g = self.tab[i]
            g.table_index = i
            g.is_119 = 1 if pid == "119" else 0
            sync = pr.get("sync")
            g.is_ms = 1 if isinstance(sync, list) and sync else 0
            if "clockabs" in pr:
                g.has_clockabs = 1
                g.clockabs = float(pr["clockabs"])
            for name, key in (("sync", "sync"), ("start", "start"), ("one", "one"), ("zero", "zero"),
                              ("flt", "float"), ("pause", "pause"), ("end", "end")):
                vals = _numlist(pr.get(key))[:GEN_MAXLIST]
                arr = getattr(g, name)
                for k, v in enumerate(vals):
                    arr[k] = v
                setattr(g, "n" + ("float" if name == "flt" else name), len(vals))
            g.length_min = int(pr["length_min"]) if pr.get("length_min") is not None else -1
            g.length_max = int(pr["length_max"]) if pr.get("length_max") is not None else -1
            g.reconstruct = 1 if pr.get("reconstructBit") else 0
            meth = pr.get("method")
            g.method = METHOD_IDS.get(meth.split(".")[-1], 0) if meth else 0
            cr = pr.get("clockrange")
            if cr and len(cr) >= 2:
                g.clock_min, g.clock_max = int(cr[0]), int(cr[1])
            g.polarity_invert = 1 if pr.get("polarity") == "invert" else 0
            if g.is_ms:
                ms_ids.append(i)
            if g.has_clockabs and g.none > 0:
                mu_ids.extend([i] * (1 if g.is_ms else 3))      # MU-only ids weighted 3:1
        self.ms_ids = np.asarray(ms_ids, dtype=np.int32)
        self.mu_ids = np.asarray(mu_ids, dtype=np.int32)
        self.mc_ids = np.asarray([i for i in range(n) if 1 <= self.tab[i].method <= 13], dtype=np.int32)
        self.mn_ids = np.asarray([i for i in range(n) if self.tab[i].method >= 14], dtype=np.int32)

    def pulse(self, kind: int, n: int, seed: int | None = None, lo: int = 0, hi: int | None = None) -> pack.PulseBatch:
        """Messages [lo, hi) of an n-message MS (kind 0) or MU (kind 1) corpus."""
        if hi is None:
            hi = n
        if seed is None:
            seed = SEED_MS if kind == pack.KIND_MS else SEED_MU
        cnt = hi - lo
        ids = self.ms_ids if kind == pack.KIND_MS else self.mu_ids
        msgs = np.zeros(cnt, dtype=pack.PULSE_DTYPE)
        rssi = np.zeros(cnt, dtype=np.int16)
        pool_p = C.c_void_p()
        pool_n = C.c_int64()
        rc = lib().sdc_gen_pulse(self.tab, C.c_int(len(self.ids)), C.c_void_p(ids.ctypes.data), C.c_int(len(ids)),
                                 C.c_int(kind), C.c_uint64(seed), C.c_int64(lo), C.c_int64(hi),
                                 C.c_void_p(msgs.ctypes.data), C.c_void_p(rssi.ctypes.data),
                                 C.byref(pool_p), C.byref(pool_n))
        if rc != 0:
            raise RuntimeError("corpus generation failed")
        digits = np.ctypeslib.as_array(C.cast(pool_p, C.POINTER(C.c_uint8)), shape=(pool_n.value,)).copy()
        lib().sdc_free(pool_p)
        clock = np.zeros(cnt, dtype=np.float64)
        if kind == pack.KIND_MS:
            cp = msgs["cp"].astype(np.int64)
            ok = cp != 0xFF
            clock[ok] = np.abs(msgs["pat"][np.nonzero(ok)[0], cp[ok]]).astype(np.float64)
        return pack.PulseBatch(kind, msgs, digits, RssiCodes(rssi), clock)


    def render_lines(self, batch: pack.PulseBatch, framed: bool = False):
        """Payload lines of a packed MS / MU batch: (text uint8, line_off uint32, line_len uint32), '\\n'-separated;
        framed=True wraps every line in STX / ETX (what the serial transport delivers)."""
        L = lib()
        L.sdc_render_lines.restype = C.c_int64
        n = batch.n
        rssi = np.ascontiguousarray(batch.rssi.codes if isinstance(batch.rssi, RssiCodes) else np.full(n, -1), dtype=np.int16)
        msgs = np.ascontiguousarray(batch.msgs)
        digits = np.ascontiguousarray(batch.digits)
        off = np.zeros(n, dtype=np.uint32)
        ln = np.zeros(n, dtype=np.uint32)
        cap = int(64 * n + 2 * int(batch.msgs["dlen"].astype(np.int64).sum()) + 64)
        while True:
            text = np.empty(cap, dtype=np.uint8)
            used = L.sdc_render_lines(C.c_int(batch.kind), C.c_void_p(msgs.ctypes.data), C.c_void_p(digits.ctypes.data),
                                      C.c_void_p(rssi.ctypes.data), C.c_int64(n), C.c_void_p(text.ctypes.data), C.c_int64(cap),
                                      C.c_void_p(off.ctypes.data), C.c_void_p(ln.ctypes.data), C.c_int(1 if framed else 0))
            if used <= cap:
                return text[:used], off, ln
            cap = int(used) + 64

    def hexmsgs(self, kind: int, n: int, seed: int | None = None, lo: int = 0, hi: int | None = None) -> pack.HexBatch:
        """Messages [lo, hi) of an n-message MC (kind 2) or MN (kind 3) corpus."""
        if hi is None:
            hi = n
        if seed is None:
            seed = SEED_MC if kind == pack.KIND_MC else SEED_MC + 0x100
        cnt = hi - lo
        ids = self.mc_ids if kind == pack.KIND_MC else self.mn_ids
        osv = -1
        if kind == pack.KIND_MC and "10" in self.ids:
            pos = np.nonzero(ids == self.ids.index("10"))[0]
            osv = int(pos[0]) if len(pos) else -1
        msgs = np.zeros(cnt, dtype=pack.HEX_DTYPE)
        pool_p = C.c_void_p()
        pool_n = C.c_int64()
        L = lib()
        L.sdc_gen_hex.restype = C.c_int
        rc = L.sdc_gen_hex(self.tab, C.c_int(len(self.ids)), C.c_void_p(ids.ctypes.data), C.c_int(len(ids)), C.c_int(osv),
                           C.c_int(kind), C.c_uint64(seed), C.c_int64(lo), C.c_int64(hi),
                           C.c_void_p(msgs.ctypes.data), C.byref(pool_p), C.byref(pool_n))
        if rc != 0:
            raise RuntimeError("corpus generation failed")
        digits = np.ctypeslib.as_array(C.cast(pool_p, C.POINTER(C.c_uint8)), shape=(pool_n.value,)).copy()
        L.sdc_free(pool_p)
        return pack.HexBatch(kind, msgs, digits, _LazyIds(self.ids, msgs["proto"]), None)


class _LazyIds:
    """protocol_id per message of a generated HexBatch (looked up on demand)."""

    def __init__(self, ids, proto):
        self.ids, self.proto = ids, proto

    def __len__(self):
        return len(self.proto)

    def __getitem__(self, i):
        return self.ids[int(self.proto[i])]


def batch_to_dicts(batch) -> List[Dict[str, Any]]:
    """Parser-style dicts for the CPU reference (same messages, same slot order)."""
    if isinstance(batch, pack.HexBatch):
        return [pack.unpack_hex(batch, i) for i in range(batch.n)]
    return [pack.unpack_pulse(batch, i) for i in range(batch.n)]
