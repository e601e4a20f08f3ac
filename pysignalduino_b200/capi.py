"""ctypes binding of libsdb200.so (include/sdb200.h) — the only way the package reaches the GPU.

There is no CPU fallback: if the library is missing it is built with nvcc, and if that or
``sdb_create`` fails (no CUDA device) the error propagates to the caller.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path
from typing import List, Optional, Tuple

import numpy as np

from . import pack
from .table import CompiledTable

import os

CHECKED = os.environ.get("SDB200_CHECKED") == "1"      # use the bounds-check build (libsdb200_chk.so)
LIB_PATH = Path(__file__).resolve().parent / ("libsdb200_chk.so" if CHECKED else "libsdb200.so")
if os.environ.get("SDB200_LIB"):                       # an experiment variant built with build_ext.py -D... --out=...
    LIB_PATH = Path(os.environ["SDB200_LIB"]).resolve()

SDB_OK, SDB_E_ARG, SDB_E_CUDA, SDB_E_OVERFLOW, SDB_E_NOGPU, SDB_E_SCRATCH = 0, -1, -2, -3, -4, -5
ST_OK, ST_INDEXERROR, ST_TYPEERROR, ST_VALUEERROR, ST_DOMAIN, ST_SCRATCH = 0, 1, 2, 3, 4, 5
STATUS_EXC = {ST_INDEXERROR: IndexError, ST_TYPEERROR: TypeError, ST_VALUEERROR: ValueError, ST_DOMAIN: pack.DomainError}
STATUS_NAMES = {ST_OK: "ok", ST_INDEXERROR: "IndexError", ST_TYPEERROR: "TypeError", ST_VALUEERROR: "ValueError",
                ST_DOMAIN: "DomainError",    # DomainError: outside the packed domain, NOT decoded (never a reference outcome)
                ST_SCRATCH: "ScratchShort"}  # device-pointer calls only: not decoded, scratch too small (Engine.scratch_short())

HIT_HAS_F, HIT_LIST, HIT_FIELDS, HIT_MM_HOST = 0x01, 0x02, 0x04, 0x08

EXPORTS = [
    "sdb_abi_version", "sdb_last_error", "sdb_create", "sdb_destroy",
    "sdb_demod_pulse_device", "sdb_demod_hex_device", "sdb_demod_host",
    "sdb_format_hits", "sdb_unit_postdemod", "sdb_unit_mc", "sdb_debug_violations", "sdb_demod_lines_host", "sdb_format_json", "sdb_frame_lines", "sdb_frame_lines_inplace",
    "sdb_unit_pattern_exists", "sdb_demod_host_payloads", "sdb_reserve", "sdb_scratch_short", "sdb_scratch_budget", "sdb_scratch_info",
]


class SdbError(RuntimeError):
    pass


class NoGpuError(SdbError):
    """No CUDA device: the demodulator has no CPU path."""


_lib: Optional[C.CDLL] = None


def load_library() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        from .build_ext import build

        build(check=CHECKED)
    L = C.CDLL(str(LIB_PATH))
    L.sdb_abi_version.restype = C.c_int
    L.sdb_last_error.restype = C.c_char_p
    L.sdb_last_error.argtypes = [C.c_void_p]
    L.sdb_create.restype = C.c_int
    L.sdb_create.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.POINTER(C.c_void_p)]
    L.sdb_destroy.restype = None
    L.sdb_destroy.argtypes = [C.c_void_p]
    dev_args = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32,
                C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
    L.sdb_demod_pulse_device.restype = C.c_int
    L.sdb_demod_pulse_device.argtypes = dev_args
    L.sdb_demod_hex_device.restype = C.c_int
    L.sdb_demod_hex_device.argtypes = [C.c_void_p, C.c_int, C.c_int] + dev_args[2:]
    L.sdb_demod_host.restype = C.c_int
    L.sdb_demod_host.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32,
                                 C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p]
    L.sdb_demod_host_payloads.restype = C.c_int
    L.sdb_demod_host_payloads.argtypes = list(L.sdb_demod_host.argtypes) + [C.c_void_p, C.c_size_t, C.c_void_p, C.POINTER(C.c_size_t)]
    L.sdb_reserve.restype = C.c_int
    L.sdb_reserve.argtypes = [C.c_void_p, C.c_uint32]
    L.sdb_scratch_short.restype = C.c_int
    L.sdb_scratch_short.argtypes = [C.c_void_p, C.POINTER(C.c_uint32)]
    L.sdb_scratch_budget.restype = C.c_int
    L.sdb_scratch_budget.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
    L.sdb_scratch_info.restype = C.c_size_t
    L.sdb_scratch_info.argtypes = [C.c_void_p, C.POINTER(C.c_uint32 * 8)]
    L.sdb_format_hits.restype = C.c_int
    L.sdb_format_hits.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_size_t,
                                  C.c_void_p, C.POINTER(C.c_size_t)]
    L.sdb_unit_postdemod.restype = C.c_int
    L.sdb_unit_postdemod.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32,
                                     C.POINTER(C.c_uint32), C.POINTER(C.c_int)]
    L.sdb_unit_mc.restype = C.c_int
    L.sdb_unit_mc.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_void_p, C.c_uint32, C.c_int, C.c_void_p, C.c_uint32,
                              C.POINTER(C.c_uint32), C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32),
                              C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.sdb_demod_lines_host.restype = C.c_int
    L.sdb_demod_lines_host.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_uint32,
                                       C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
    L.sdb_format_json.restype = C.c_int
    L.sdb_format_json.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint32, C.c_void_p, C.c_char_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.POINTER(C.c_size_t)]
    L.sdb_frame_lines.restype = C.c_int
    L.sdb_frame_lines.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32,
                                  C.POINTER(C.c_uint32), C.POINTER(C.c_size_t)]
    L.sdb_frame_lines_inplace.restype = C.c_int
    L.sdb_frame_lines_inplace.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p,
                                          C.c_size_t, C.POINTER(C.c_uint32), C.POINTER(C.c_size_t)]
    L.sdb_unit_pattern_exists.restype = C.c_int
    L.sdb_unit_pattern_exists.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32,
                                          C.c_void_p, C.c_size_t, C.c_uint32, C.POINTER(C.c_int), C.c_void_p, C.c_uint32,
                                          C.POINTER(C.c_int)]
    L.sdb_debug_violations.restype = C.c_uint
    L.sdb_debug_violations.argtypes = [C.c_void_p, C.c_int]
    if L.sdb_abi_version() != 3:
        raise SdbError("libsdb200.so ABI version mismatch")
    _lib = L
    return L


LINEINFO_DTYPE = np.dtype([("status", "u1"), ("has_r", "u1"), ("r_len", "<u2"), ("r_off", "<u4"), ("clock", "<i4")])
LINE_INVALID, LINE_OK, LINE_HOSTPATH = 0, 1, 2


FRAME_OTHER, FRAME_SIDE, FRAME_PYPATH, FRAME_NONE = 4, 0x40, 0x80, 0xFF


def frame_lines_inplace(raw: bytes):
    """Raw receive buffer -> (line_off, line_len, line_type, side text): plain payloads are addressed inside ``raw``,
    reduced ones are decompressed into the side buffer (type flag FRAME_SIDE)."""
    L = load_library()
    buf = np.frombuffer(raw, dtype=np.uint8)
    raw = buf                                  # (bytes, memoryview and arrays alike)
    max_lines = len(raw) // 48 + 1024          # a guess; the call reports what it needs (bytes.count would cost more than the framing)
    cap = 4096
    while True:
        side = np.empty(cap, dtype=np.uint8)
        off = np.empty(max_lines, dtype=np.uint32)
        ln = np.empty(max_lines, dtype=np.uint32)
        typ = np.empty(max_lines, dtype=np.uint8)
        n, used = C.c_uint32(0), C.c_size_t(0)
        rc = L.sdb_frame_lines_inplace(buf.ctypes.data if len(buf) else None, len(buf), off.ctypes.data, ln.ctypes.data,
                                       typ.ctypes.data, max_lines, side.ctypes.data, cap, C.byref(n), C.byref(used))
        if rc == SDB_E_OVERFLOW:
            max_lines, cap = max(max_lines, n.value), max(cap, used.value + 64)
            continue
        if rc != SDB_OK:
            raise SdbError(f"sdb_frame_lines_inplace failed ({rc})")
        k = n.value
        return off[:k], ln[:k], typ[:k], side[: used.value]


def frame_lines(raw: bytes):
    """Raw receive buffer -> (payload text uint8, line_off, line_len, line_type) per '\\n'-separated raw line
    (base.py:13-193 natively; needs no GPU)."""
    L = load_library()
    buf = np.frombuffer(raw, dtype=np.uint8)
    max_lines = raw.count(b"\n") + 1
    cap = 2 * len(raw) + 64
    while True:
        text = np.empty(cap, dtype=np.uint8)
        off = np.zeros(max_lines, dtype=np.uint32)
        ln = np.zeros(max_lines, dtype=np.uint32)
        typ = np.zeros(max_lines, dtype=np.uint8)
        n, used = C.c_uint32(0), C.c_size_t(0)
        rc = L.sdb_frame_lines(buf.ctypes.data if len(buf) else None, len(buf), text.ctypes.data, cap, off.ctypes.data,
                               ln.ctypes.data, typ.ctypes.data, max_lines, C.byref(n), C.byref(used))
        if rc == SDB_E_OVERFLOW:
            max_lines, cap = max(max_lines, n.value), max(cap, used.value + 64)
            continue
        if rc != SDB_OK:
            raise SdbError(f"sdb_frame_lines failed ({rc})")
        k = n.value
        return text[: used.value], off[:k], ln[:k], typ[:k]


def frame_chunks(raw, chunk_bytes: int = 256 << 20):
    """Frame a large receive buffer piecewise, one chunk ahead of the consumer: yields ``(byte_base, byte_end, line_base,
    off, ln, typ, side)`` per chunk (``off`` relative to ``raw[byte_base:byte_end]`` for plain payloads).  The framing of
    chunk k + 1 runs on the host threads (ctypes drops the GIL) while the caller feeds chunk k to the device.  Chunks end
    on a line boundary and stay below 4 GiB (the device path addresses text with 32-bit offsets)."""
    from concurrent.futures import ThreadPoolExecutor

    view = memoryview(raw)
    n = len(view)
    chunk_bytes = max(1 << 16, min(chunk_bytes, 1 << 31))
    cuts = [0]
    while cuts[-1] < n:
        pos = cuts[-1] + chunk_bytes
        if pos >= n:
            cuts.append(n)
            break
        cut = n
        probe = pos
        while probe < n:                                            # lines are far shorter than 1 MiB; keep looking if not
            if probe - cuts[-1] >= (1 << 32) - (2 << 20):
                raise SdbError("frame_chunks: a single line of more than 3.9 GiB cannot be framed")
            nl = bytes(view[probe : min(n, probe + (1 << 20))]).find(b"\n")
            if nl >= 0:
                cut = probe + nl + 1
                break
            probe += 1 << 20
        cuts.append(cut)
    with ThreadPoolExecutor(max_workers=1) as ex:
        fut = ex.submit(frame_lines_inplace, view[cuts[0] : cuts[1]]) if len(cuts) > 1 else None
        line_base = 0
        for k in range(len(cuts) - 1):
            off, ln, typ, side = fut.result()
            fut = ex.submit(frame_lines_inplace, view[cuts[k + 1] : cuts[k + 2]]) if k + 2 < len(cuts) else None
            yield cuts[k], cuts[k + 1], line_base, off, ln, typ, side
            line_base += len(typ)


def bind_to_gpu_numa(device: int = 0):
    """Pin the calling process to the CPU cores next to CUDA device ``device`` (NVML's ideal CPU affinity) BEFORE it allocates
    pinned host buffers: pinned pages are placed on the NUMA node of the thread that first touches them, and a rank whose
    buffers sit on the other socket pushes every H2D / D2H byte over the inter-socket link.  With one process per GPU on a
    multi-socket host this is what keeps the end-to-end path scaling.  Returns the CPU list used (None if NVML or the
    affinity call is unavailable — the call is then a no-op)."""
    try:
        import pynvml

        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        index = int(vis.split(",")[device]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else device
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [64 * w + b for w, x in enumerate(words) for b in range(64) if (int(x) >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:                      # no NVML / not permitted: keep the inherited affinity
        return None


class Result:
    """Raw result arrays of one batch call."""

    __slots__ = ("kind", "out", "hits", "bits", "counters")

    def __init__(self, kind, out, hits, bits, counters):
        self.kind, self.out, self.hits, self.bits, self.counters = kind, out, hits, bits, counters


class Engine:
    """One sdb handle = one compiled protocol table resident on one GPU."""

    def __init__(self, table: CompiledTable, device: int = 0):
        self.lib = load_library()
        self.table = table
        self.device = device
        h = C.c_void_p()
        blob = table.blob
        rc = self.lib.sdb_create(blob, len(blob), device, C.byref(h))
        if rc != SDB_OK:
            msg = (self.lib.sdb_last_error(None) or b"").decode()
            raise (NoGpuError if rc == SDB_E_NOGPU else SdbError)(f"sdb_create failed ({rc}): {msg}")
        self.h = h

    def close(self) -> None:
        if getattr(self, "h", None):
            self.lib.sdb_destroy(self.h)
            self.h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def reserve(self, n_messages: int) -> None:
        """Pre-size the scratch so that later device-pointer calls only enqueue work (sdb_reserve)."""
        rc = self.lib.sdb_reserve(self.h, n_messages)
        if rc != SDB_OK:
            raise self._err(rc, "sdb_reserve")

    def scratch_short(self) -> int:
        """After device-pointer calls: synchronise, return how many messages were flagged ST_SCRATCH since the last check
        (their launch group needed more scratch than the compact arenas hold) and grow the scratch budgets so that
        submitting those messages again succeeds (sdb_scratch_short).  The host-buffer calls do this themselves."""
        n = C.c_uint32(0)
        rc = self.lib.sdb_scratch_short(self.h, C.byref(n))
        if rc != SDB_OK:
            raise self._err(rc, "sdb_scratch_short")
        return int(n.value)

    def scratch_budget(self, surv_avg: int = 0, match_avg: int = 0, ovf_max: int = 0, slack_warps: int = 0) -> None:
        """Budgets the scratch is sized by (0 = keep): survivor / MU match records per message on average, messages of the
        worst-case overflow region; ``slack_warps`` caps the arena blocks added on top (tests only) (sdb_scratch_budget)."""
        rc = self.lib.sdb_scratch_budget(self.h, surv_avg, match_avg, ovf_max, slack_warps)
        if rc != SDB_OK:
            raise self._err(rc, "sdb_scratch_budget")

    def scratch_info(self) -> dict:
        """Size and budgets of the scratch block, and the need the last ``scratch_short()`` / host-buffer call read back."""
        cfg = (C.c_uint32 * 8)()
        nbytes = self.lib.sdb_scratch_info(self.h, C.byref(cfg))
        return {"bytes": int(nbytes), "chunk": int(cfg[0]), "surv_avg": int(cfg[1]), "match_avg": int(cfg[2]), "ovf_max": int(cfg[3]),
                "arena_blocks": int(cfg[4]), "need_surv_records": int(cfg[5]), "need_match_records": int(cfg[6]),
                "need_overflow_messages": int(cfg[7])}

    def _err(self, rc: int, what: str) -> SdbError:
        return SdbError(f"{what} failed ({rc}): {(self.lib.sdb_last_error(self.h) or b'').decode()}")

    # ---- host-buffer call ------------------------------------------------------------------
    def demod_host(self, batch, mc_repaired: bool = False, hits_cap: int = 0, bits_cap: int = 0) -> Result:
        n = batch.n
        kind = batch.kind
        msgs = np.ascontiguousarray(batch.msgs)
        digits = np.ascontiguousarray(batch.digits)
        out = np.zeros(n, dtype=pack.MSGOUT_DTYPE)
        hits_cap = hits_cap or max(1024, 4 * n)
        bits_cap = bits_cap or max(4096, 16 * n)
        ctr = np.zeros(1, dtype=pack.COUNTERS_DTYPE)
        while True:
            hits = np.empty(hits_cap, dtype=pack.HIT_DTYPE)
            bits = np.empty(bits_cap, dtype=np.uint32)
            rc = self.lib.sdb_demod_host(self.h, kind, 1 if mc_repaired else 0, msgs.ctypes.data, digits.ctypes.data,
                                         digits.nbytes, n, out.ctypes.data, hits.ctypes.data, hits_cap,
                                         bits.ctypes.data, bits_cap, ctr.ctypes.data)
            if rc == SDB_E_OVERFLOW:
                hits_cap = max(hits_cap, int(ctr["hits"][0]) + 16)
                bits_cap = max(bits_cap, int(ctr["words"][0]) + 16)
                continue
            if rc != SDB_OK:
                raise self._err(rc, "sdb_demod_host")
            nh, nw = int(ctr["hits"][0]), int(ctr["words"][0])
            return Result(kind, out, hits[:nh], bits[:nw], ctr[0])

    def demod_host_into(self, kind: int, msgs: np.ndarray, digits: np.ndarray, out: np.ndarray, hits: np.ndarray,
                        bits: np.ndarray, ctr: np.ndarray, mc_repaired: bool = False) -> int:
        """Same call with caller-owned (e.g. pinned) host arrays; returns the raw code (SDB_OK / SDB_E_OVERFLOW)."""
        rc = self.lib.sdb_demod_host(self.h, kind, 1 if mc_repaired else 0, msgs.ctypes.data, digits.ctypes.data,
                                     digits.nbytes, len(msgs), out.ctypes.data, hits.ctypes.data, len(hits),
                                     bits.ctypes.data, len(bits), ctr.ctypes.data)
        if rc not in (SDB_OK, SDB_E_OVERFLOW):
            raise self._err(rc, "sdb_demod_host")
        return rc

    def demod_host_payloads_into(self, kind: int, msgs: np.ndarray, digits: np.ndarray, out: np.ndarray, phits: np.ndarray,
                                 ctr: np.ndarray, pool: np.ndarray, mc_repaired: bool = False, bits_cap: int = 0,
                                 hits: Optional[np.ndarray] = None, bits: Optional[np.ndarray] = None):
        """Decode + payload strings in one pipelined call with caller-owned (ideally pinned) arrays: ``phits`` is a
        ``pack.PAYHIT_DTYPE`` array (12 bytes per hit: string offset, protocol, bit length, flags), ``pool`` uint8; hit i's
        string starts at ``pool[phits[i]["off"]]`` and is NUL-terminated.  The 16-byte ``hits`` records and the ``bits``
        arena are optional extras (None: they stay on the device; ``bits_cap`` sizes the device arena, default 16 words per
        hit slot).  Returns (raw code, pool bytes used)."""
        used = C.c_size_t(0)
        if hits is not None and len(hits) < len(phits):
            raise ValueError("hits must have at least len(phits) entries")
        rc = self.lib.sdb_demod_host_payloads(self.h, kind, 1 if mc_repaired else 0, msgs.ctypes.data, digits.ctypes.data,
                                              digits.nbytes, len(msgs), out.ctypes.data,
                                              hits.ctypes.data if hits is not None else None, len(phits),
                                              bits.ctypes.data if bits is not None else None,
                                              len(bits) if bits is not None else (bits_cap or max(4096, 16 * len(phits))), ctr.ctypes.data,
                                              pool.ctypes.data, pool.nbytes, phits.ctypes.data, C.byref(used))
        if rc not in (SDB_OK, SDB_E_OVERFLOW):
            raise self._err(rc, "sdb_demod_host_payloads")
        return rc, used.value

    def demod_payloads(self, batch, mc_repaired: bool = False):
        """Convenience form: -> (Result whose ``hits`` are PAYHIT_DTYPE records and whose ``bits`` are empty, pool bytes)."""
        n = batch.n
        kind = batch.kind
        msgs = np.ascontiguousarray(batch.msgs)
        digits = np.ascontiguousarray(batch.digits)
        out = np.zeros(n, dtype=pack.MSGOUT_DTYPE)
        ctr = np.zeros(1, dtype=pack.COUNTERS_DTYPE)
        hits_cap, bits_cap, pool_cap = max(1024, 4 * n), max(4096, 16 * n), max(4096, 96 * n)
        while True:
            phits = np.empty(hits_cap, dtype=pack.PAYHIT_DTYPE)
            pool = np.empty(pool_cap, dtype=np.uint8)
            rc, used = self.demod_host_payloads_into(kind, msgs, digits, out, phits, ctr, pool, mc_repaired=mc_repaired, bits_cap=bits_cap)
            if rc == SDB_E_OVERFLOW:
                hits_cap = max(hits_cap, int(ctr["hits"][0]) + 16)
                bits_cap = max(bits_cap, int(ctr["words"][0]) + 16)
                pool_cap = max(pool_cap, used + 64)
                continue
            nh = int(ctr["hits"][0])
            return Result(kind, out, phits[:nh], np.zeros(0, dtype=np.uint32), ctr[0]), pool[:used]

    # ---- text lines (tokenizer kernel + demodulation) ---------------------------------------
    def demod_lines(self, kind: int, text: np.ndarray, line_off: np.ndarray, line_len: np.ndarray,
                    hits_cap: int = 0, bits_cap: int = 0):
        """Payload lines of one type (MS / MU) -> (Result indexed by line, SdbLineInfo array)."""
        n = len(line_off)
        text = np.ascontiguousarray(text, dtype=np.uint8)
        line_off = np.ascontiguousarray(line_off, dtype=np.uint32)
        line_len = np.ascontiguousarray(line_len, dtype=np.uint32)
        out = np.zeros(n, dtype=pack.MSGOUT_DTYPE)
        info = np.zeros(n, dtype=LINEINFO_DTYPE)
        hits_cap = hits_cap or max(1024, 4 * n)
        bits_cap = bits_cap or max(4096, 16 * n)
        ctr = np.zeros(1, dtype=pack.COUNTERS_DTYPE)
        while True:
            hits = np.empty(hits_cap, dtype=pack.HIT_DTYPE)
            bits = np.empty(bits_cap, dtype=np.uint32)
            rc = self.lib.sdb_demod_lines_host(self.h, kind, text.ctypes.data, text.nbytes, line_off.ctypes.data,
                                               line_len.ctypes.data, n, out.ctypes.data, hits.ctypes.data, hits_cap,
                                               bits.ctypes.data, bits_cap, ctr.ctypes.data, info.ctypes.data)
            if rc == SDB_E_OVERFLOW:
                hits_cap = max(hits_cap, int(ctr["hits"][0]) + 16)
                bits_cap = max(bits_cap, int(ctr["words"][0]) + 16)
                continue
            if rc != SDB_OK:
                raise self._err(rc, "sdb_demod_lines_host")
            nh, nw = int(ctr["hits"][0]), int(ctr["words"][0])
            return Result(kind, out, hits[:nh], bits[:nw], ctr[0]), info

    # ---- device-pointer call (pointers are ints, e.g. torch tensor.data_ptr()) ---------------
    def demod_pulse_device(self, kind: int, d_msgs: int, d_digits: int, n: int, d_out: int, d_hits: int, hits_cap: int,
                           d_bits: int, bits_cap: int, d_ctr: int, stream: int = 0) -> None:
        rc = self.lib.sdb_demod_pulse_device(self.h, kind, d_msgs, d_digits, n, d_out, d_hits, hits_cap, d_bits,
                                             bits_cap, d_ctr, stream)
        if rc != SDB_OK:
            raise self._err(rc, "sdb_demod_pulse_device")

    def demod_hex_device(self, kind: int, mc_repaired: bool, d_msgs: int, d_digits: int, n: int, d_out: int,
                         d_hits: int, hits_cap: int, d_bits: int, bits_cap: int, d_ctr: int, stream: int = 0) -> None:
        rc = self.lib.sdb_demod_hex_device(self.h, kind, 1 if mc_repaired else 0, d_msgs, d_digits, n, d_out, d_hits,
                                           hits_cap, d_bits, bits_cap, d_ctr, stream)
        if rc != SDB_OK:
            raise self._err(rc, "sdb_demod_hex_device")

    # ---- formatting ------------------------------------------------------------------------
    def format_hits(self, kind: int, hits: np.ndarray, bits: np.ndarray) -> Tuple[bytes, np.ndarray]:
        """Payload strings of all hits: (pool bytes, offsets[nhits+1])."""
        nh = len(hits)
        hits = np.ascontiguousarray(hits)
        bits = np.ascontiguousarray(bits)
        off = np.zeros(nh + 1, dtype=np.uint64)
        cap = max(64, 40 * nh)
        while True:
            pool = np.empty(cap, dtype=np.uint8)
            used = C.c_size_t(0)
            rc = self.lib.sdb_format_hits(self.h, kind, hits.ctypes.data, nh, bits.ctypes.data if len(bits) else None,
                                          pool.ctypes.data, cap, off.ctypes.data, C.byref(used))
            if rc == SDB_E_OVERFLOW:
                cap = used.value + 16
                continue
            if rc != SDB_OK:
                raise self._err(rc, "sdb_format_hits")
            return pool[: used.value].tobytes(), off

    def format_hits_into(self, kind: int, hits: np.ndarray, bits: np.ndarray, pool: np.ndarray, off: np.ndarray) -> int:
        """Same with caller-owned buffers (``pool`` uint8, ``off`` uint64[len(hits) + 1]); returns the bytes used
        (raises when the pool is too small).  The formatter splits large batches over the host threads."""
        used = C.c_size_t(0)
        rc = self.lib.sdb_format_hits(self.h, kind, hits.ctypes.data if len(hits) else None, len(hits),
                                      bits.ctypes.data if len(bits) else None, pool.ctypes.data, pool.nbytes, off.ctypes.data,
                                      C.byref(used))
        if rc != SDB_OK:
            raise self._err(rc, "sdb_format_hits")
        return used.value

    def format_json(self, kind: int, hits: np.ndarray, bits: np.ndarray, text: np.ndarray, line_off: np.ndarray,
                    info: np.ndarray) -> Tuple[bytes, np.ndarray]:
        """MqttPublisher._message_to_json of every MS / MU hit of a demod_lines call: (pool bytes, offsets[nhits+1])."""
        nh = len(hits)
        hits = np.ascontiguousarray(hits)
        bits = np.ascontiguousarray(bits)
        text = np.ascontiguousarray(text, dtype=np.uint8)
        line_off = np.ascontiguousarray(line_off, dtype=np.uint32)
        info = np.ascontiguousarray(info)
        if not hasattr(self, "_id_pool"):
            enc = [i.encode("utf-8") for i in self.table.ids]
            self._id_pool = b"".join(enc) + b"\0"
            self._id_off = np.zeros(len(enc) + 1, dtype=np.uint32)
            np.cumsum([len(e) for e in enc], out=self._id_off[1:])
        off = np.zeros(nh + 1, dtype=np.uint64)
        cap = max(256, 200 * nh)
        while True:
            pool = np.empty(cap, dtype=np.uint8)
            used = C.c_size_t(0)
            rc = self.lib.sdb_format_json(self.h, kind, hits.ctypes.data, nh, bits.ctypes.data if len(bits) else None,
                                          self._id_pool, self._id_off.ctypes.data, text.ctypes.data, line_off.ctypes.data,
                                          info.ctypes.data, pool.ctypes.data, cap, off.ctypes.data, C.byref(used))
            if rc == SDB_E_OVERFLOW:
                cap = used.value + 16
                continue
            if rc != SDB_OK:
                raise self._err(rc, "sdb_format_json")
            return pool[: used.value].tobytes(), off

    def unit_mc(self, proto_index: int, bits: str, mcbitnum: int, method_override: int = 0):
        """One mcBit2* call on the device: -> (rcode, reason, [bit strings])."""
        a = np.frombuffer(bits.encode("ascii"), dtype=np.uint8) - ord("0")
        a = np.ascontiguousarray(a.astype(np.uint8))
        out = np.zeros(4096, dtype=np.uint8)
        seg = np.zeros(64, dtype=np.int32)
        n_out, n_seg, rcode, reason = C.c_uint32(0), C.c_uint32(0), C.c_int(0), C.c_int(0)
        rc = self.lib.sdb_unit_mc(self.h, proto_index, method_override, a.ctypes.data if len(a) else None, len(a), mcbitnum,
                                  out.ctypes.data, len(out), C.byref(n_out), seg.ctypes.data, len(seg), C.byref(n_seg),
                                  C.byref(rcode), C.byref(reason))
        if rc != SDB_OK:
            raise self._err(rc, "sdb_unit_mc")
        s = "".join("01"[v] for v in out[: n_out.value])
        if n_seg.value:
            parts, o = [], 0
            for ln in seg[: n_seg.value]:
                parts.append(s[o : o + int(ln)])
                o += int(ln)
            return rcode.value, reason.value, parts
        return rcode.value, reason.value, [s]

    def unit_pattern_exists(self, tpl: np.ndarray, rank: np.ndarray, tenths: np.ndarray, pat_ids: int, npat: int,
                            digits: np.ndarray, dlen: int):
        """One pattern_exists call on the device: -> (found, [id digit per template position], first position)."""
        tpl = np.ascontiguousarray(tpl)
        rank = np.ascontiguousarray(rank, dtype=np.uint16)
        tenths = np.ascontiguousarray(tenths, dtype=np.int16)
        digits = np.ascontiguousarray(digits, dtype=np.uint8)
        tgt = np.zeros(16, dtype=np.uint8)
        found, pos = C.c_int(0), C.c_int(0)
        rc = self.lib.sdb_unit_pattern_exists(self.h, tpl.ctypes.data, tpl.nbytes, rank.ctypes.data if len(rank) else None, len(rank),
                                              tenths.ctypes.data, pat_ids, npat, digits.ctypes.data if len(digits) else None,
                                              digits.nbytes, dlen, C.byref(found), tgt.ctypes.data, len(tgt), C.byref(pos))
        if rc != SDB_OK:
            raise self._err(rc, "sdb_unit_pattern_exists")
        return bool(found.value), [int(x) for x in tgt[: int(tpl["len"].reshape(-1)[0])]], pos.value

    def debug_violations(self, reset: bool = False) -> int:
        """Out-of-range index count of the bounds-check build (0xFFFFFFFF from the normal build)."""
        return int(self.lib.sdb_debug_violations(self.h, 1 if reset else 0))

    # ---- unit ops --------------------------------------------------------------------------
    def unit_postdemod(self, method: int, bits_in) -> Tuple[int, List[int]]:
        a = np.ascontiguousarray(np.asarray(list(bits_in), dtype=np.uint8))
        out = np.zeros(len(a) + 64, dtype=np.uint8)
        n_out, rcode = C.c_uint32(0), C.c_int(0)
        rc = self.lib.sdb_unit_postdemod(self.h, method, a.ctypes.data if len(a) else None, len(a), out.ctypes.data,
                                         len(out), C.byref(n_out), C.byref(rcode))
        if rc != SDB_OK:
            raise self._err(rc, "sdb_unit_postdemod")
        return rcode.value, [int(x) for x in out[: n_out.value]]
