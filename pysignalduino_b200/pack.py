"""Host-side packing of parser dicts into the fixed-width batch arrays of the C ABI.

Replaces the per-message input handling at the top of the reference demodulators:
  * MS gates and P# parsing  — sd_protocols/message_synced.py:21-66
  * MU gate and P# parsing   — sd_protocols/message_unsynced.py:22-35
  * MC / MN argument pickup  — sd_protocols/sd_protocols.py:79-88, :115-129

Layout (include/sdb200.h): ``SdbPulseMsg`` 48 B/message + a nibble-packed digit pool whose
per-message streams start on 16-byte boundaries and are padded with 0xF nibbles.

Inputs the packed domain cannot represent (non-integer pulse values, pattern ids >= 10,
more than 8 pattern slots, D longer than SDB_MAX_DIGITS = 4096) are marked PER MESSAGE
(``SDB_MSG_DOMAIN`` -> status ``"DomainError"``, reason in ``batch.domain``): the rest of the
batch is decoded normally, and such a message is never silently decoded differently from the
reference.  ``strict=True`` raises :class:`DomainError` at the first one instead.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np

MAX_SLOTS = 8
MAX_DIGITS = 4096        # SDB_MAX_DIGITS: the fast kernels stage 1024 digits, longer messages take the long kernels
MAX_HEX = 512
DIGIT_OTHER = 0xE
DIGIT_PAD = 0xF

KIND_MS, KIND_MU, KIND_MC, KIND_MN = 0, 1, 2, 3
KIND_BY_NAME = {"MS": KIND_MS, "MU": KIND_MU, "MC": KIND_MC, "MN": KIND_MN}

MSG_VALID = 0x01
HEX_TOGGLE_POLARITY = 0x02
MSG_DOMAIN = 0x04

PULSE_DTYPE = np.dtype(
    [
        ("pat", "<i4", (MAX_SLOTS,)),
        ("doff", "<u4"),
        ("dlen", "<u2"),
        ("npat", "u1"),
        ("cp", "u1"),
        ("pat_ids", "<u4"),
        ("flags", "u1"),
        ("rsv", "u1", (3,)),
    ]
)
assert PULSE_DTYPE.itemsize == 48

HEX_DTYPE = np.dtype(
    [
        ("doff", "<u4"),
        ("hlen", "<u2"),
        ("proto", "<u2"),
        ("clock", "<i4"),
        ("bitlen", "<i2"),
        ("flags", "u1"),
        ("rsv", "u1"),
    ]
)
assert HEX_DTYPE.itemsize == 16

MSGOUT_DTYPE = np.dtype([("hit_off", "<u4"), ("nhits", "<u2"), ("status", "u1"), ("reason", "u1")])
HIT_DTYPE = np.dtype(
    [
        ("msg", "<u4"),
        ("bits_off", "<u4"),
        ("proto", "<u2"),
        ("nbits", "<u2"),
        ("aux", "<u2"),
        ("flags", "u1"),
        ("rsv", "u1"),
    ]
)
# SdbPayloadHit: what sdb_demod_host_payloads returns per hit (the string lives in the pool at `off`)
PAYHIT_DTYPE = np.dtype([("off", "<u4"), ("proto", "<u2"), ("nbits", "<u2"), ("aux", "<u2"), ("flags", "u1"), ("rsv", "u1")])
assert PAYHIT_DTYPE.itemsize == 12
COUNTERS_DTYPE = np.dtype([("hits", "<u4"), ("words", "<u4"), ("raised", "<u4"), ("domain", "<u4")])
assert MSGOUT_DTYPE.itemsize == 8 and HIT_DTYPE.itemsize == 16 and COUNTERS_DTYPE.itemsize == 16


class DomainError(ValueError):
    """The message cannot be represented in the packed batch format."""


_DIGIT_LUT = np.full(256, DIGIT_OTHER, dtype=np.uint8)
_DIGIT_LUT[ord("0") : ord("9") + 1] = np.arange(10, dtype=np.uint8)

_HEX_LUT = np.full(256, 0xFF, dtype=np.uint8)
_HEX_LUT[ord("0") : ord("9") + 1] = np.arange(10, dtype=np.uint8)
_HEX_LUT[ord("A") : ord("F") + 1] = np.arange(10, 16, dtype=np.uint8)
_HEX_LUT[ord("a") : ord("f") + 1] = np.arange(10, 16, dtype=np.uint8)


def parse_patterns(msg_data: Dict[str, Any]) -> Dict[str, float]:
    """``P<d>`` keys -> ``{str(int(d)): float(value)}`` in dict insertion order.

    message_synced.py:50-57 / message_unsynced.py:28-35: a later duplicate id overwrites the
    value but keeps the first position; a value ``float()`` rejects is skipped.
    """
    patterns: Dict[str, float] = {}
    for key, val in msg_data.items():
        if key.startswith("P") and key[1:].isdigit():
            try:
                pidx = str(int(key[1:]))
                patterns[pidx] = float(val)
            except ValueError:
                pass
    return patterns


def _ms_gates(msg_data: Dict[str, Any]) -> bool:
    """message_synced.py:21-47 — D, CP, SP (and R when present) must be digit strings."""
    raw_data = msg_data.get("data", "")
    if not raw_data or not raw_data.isdigit():
        return False
    cp = msg_data.get("CP", "")
    if not cp or not cp.isdigit():
        return False
    sp = msg_data.get("SP", "")
    if not sp or not sp.isdigit():
        return False
    if "R" in msg_data:
        if not msg_data.get("R", "").isdigit():
            return False
    return True


class PulseBatch:
    """Packed MS or MU batch plus the host-only fields needed to format results."""

    __slots__ = ("kind", "msgs", "digits", "rssi", "clock", "n", "domain")

    def __init__(self, kind: int, msgs: np.ndarray, digits: np.ndarray, rssi: List[Any], clock: np.ndarray,
                 domain: Optional[Dict[int, str]] = None):
        self.kind = kind
        self.msgs = msgs
        self.digits = digits
        self.rssi = rssi      # msg_data.get('R') per message (meta.rssi is the raw value)
        self.clock = clock    # MS: abs(P[CP]) per message (meta.clock); unused for MU
        self.n = len(msgs)
        self.domain = domain or {}   # message index -> why it is outside the packed domain (status "DomainError")

    def algorithmic_input_bytes(self) -> int:
        """48 B header + ceil(dlen/2) digit bytes per message."""
        return int(48 * self.n + ((self.msgs["dlen"].astype(np.int64) + 1) // 2).sum())


def pack_digit_streams(streams: Sequence[np.ndarray]) -> Tuple[np.ndarray, np.ndarray]:
    """Nibble-pack digit arrays (values 0..15) into a pool; returns (pool, doff in 16-byte units)."""
    n = len(streams)
    lens = np.fromiter((len(s) for s in streams), dtype=np.int64, count=n)
    padded = (lens + 31) // 32 * 32            # 32 nibbles = 16 bytes
    offs = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(padded, out=offs[1:])
    nib = np.full(int(offs[-1]) + 64, DIGIT_PAD, dtype=np.uint8)   # +32 B tail so device windows may over-read
    for i, s in enumerate(streams):
        if len(s):
            nib[offs[i] : offs[i] + len(s)] = s
    pool = (nib[0::2] | (nib[1::2] << 4)).astype(np.uint8)
    return pool, (offs[:-1] // 32).astype(np.uint32)


_PAD32 = [b"\xff" * k for k in range(32)]          # tail padding to the 16-byte unit (0xFF -> nibble 0xF)
_DIGIT_LUT_PAD = _DIGIT_LUT.copy()
_DIGIT_LUT_PAD[0xFF] = DIGIT_PAD


def _pulse_domain_reason(data, patterns: Dict[str, float]) -> Optional[str]:
    """Why the packed record cannot hold this (gate-passing) message, or None."""
    if not isinstance(data, str):
        return "'data' must be a str"
    if len(data) > MAX_DIGITS:
        return f"D has {len(data)} digits (max {MAX_DIGITS})"
    if len(patterns) > MAX_SLOTS:
        return f"{len(patterns)} pattern slots (max {MAX_SLOTS})"
    for pidx, val in patterns.items():
        if len(pidx) != 1:
            return f"pattern id {pidx!r} is not a single digit"
        if val != val or val in (float("inf"), float("-inf")) or val != int(val) or abs(val) > 2147483647:
            return f"pattern value {val!r} is not an int32"
    return None


try:                                    # native dict walker (csrc/sdb_fastpack.c, built by build_ext.build_fastpack)
    from . import _fastpack
except ImportError:                     # not built: the Python packer below does the same, ~30x slower
    _fastpack = None


def pack_pulse(msgs: Sequence[Dict[str, Any]], kind: int, strict: bool = False, native: bool = True) -> PulseBatch:
    """Pack MS (kind 0) or MU (kind 1) parser dicts.

    The native packer (``_fastpack.pack_pulse``) walks the dicts in C; it declines (returns None) on anything it cannot
    reproduce exactly — non-str values, non-ASCII text, pulse values outside the canonical integer syntax — and the Python
    packer below then handles the whole batch: one pass of plain-Python bookkeeping per message (dict lookups, list
    appends), every array operation once over the whole batch.  A message outside the packed domain becomes an
    SDB_MSG_DOMAIN record (status "DomainError" for that message only); ``strict`` raises instead."""
    n = len(msgs)
    if native and _fastpack is not None and n:
        rec = np.zeros(n, dtype=PULSE_DTYPE)
        clock = np.zeros(n, dtype=np.float64)
        got = _fastpack.pack_pulse(msgs if type(msgs) is list else list(msgs), kind, rec, clock)
        if got is not None:
            pool_bytes, rssi, domain = got
            if strict and domain:
                i = min(domain)
                raise DomainError(f"message {i}: {domain[i]}")
            return PulseBatch(kind, rec, np.frombuffer(pool_bytes, dtype=np.uint8), rssi, clock, domain)
    rssi: List[Any] = []
    pats: List[int] = []            # 8 values per message
    meta: List[int] = []            # dlen, npat, cp, pat_ids, flags per message
    chunks: List[bytes] = []        # digit characters, each message padded to a multiple of 32 with 0xFF
    clock = np.zeros(n, dtype=np.float64)
    domain: Dict[int, str] = {}
    zero8 = [0] * MAX_SLOTS
    is_ms = kind == KIND_MS
    for i, m in enumerate(msgs):
        rssi.append(m.get("R"))
        data = m.get("data", "")
        valid = _ms_gates(m) if is_ms else bool(data)          # message_unsynced.py:22-25
        if not valid:
            pats.extend(zero8)
            meta.extend((0, 0, 0xFF, 0, 0))
            continue
        patterns = parse_patterns(m)
        why = _pulse_domain_reason(data, patterns)
        raw = b""
        if why is None:
            raw = data.encode("ascii", "replace")  # one byte per character
            if b"\xff" in raw:                     # cannot happen after an ascii encode; keeps the pad byte unambiguous
                why = "D contains a 0xFF byte"
        if why is not None:
            if strict:
                raise DomainError(f"message {i}: {why}")
            domain[i] = why
            pats.extend(zero8)
            meta.extend((0, 0, 0xFF, 0, MSG_DOMAIN))
            continue
        ids = 0
        vals = zero8.copy()
        for s, (pidx, val) in enumerate(patterns.items()):
            vals[s] = int(val)
            ids |= int(pidx) << (4 * s)
        cp = 0xFF
        if is_ms:
            cp_key = str(int(m.get("CP", "")))     # message_synced.py:33,59
            if cp_key in patterns:
                cp = list(patterns).index(cp_key)
                clock[i] = abs(patterns[cp_key])
        pats.extend(vals)
        meta.extend((len(raw), len(patterns), cp, ids, MSG_VALID))
        chunks.append(raw)
        pad = -len(raw) & 31
        if pad:
            chunks.append(_PAD32[pad])
    rec = np.zeros(n, dtype=PULSE_DTYPE)
    if n:
        rec["pat"] = np.asarray(pats, dtype=np.int64).reshape(n, MAX_SLOTS).astype(np.int32)
        mt = np.asarray(meta, dtype=np.int64).reshape(n, 5)
        rec["dlen"], rec["npat"], rec["cp"], rec["pat_ids"], rec["flags"] = mt[:, 0], mt[:, 1], mt[:, 2], mt[:, 3], mt[:, 4]
        units = (mt[:, 0] + 31) // 32
        doff = np.zeros(n, dtype=np.int64)
        np.cumsum(units[:-1], out=doff[1:])
        rec["doff"] = doff
    blob = b"".join(chunks) + b"\xff" * 64        # +32 B tail so device windows may over-read
    nib = _DIGIT_LUT_PAD[np.frombuffer(blob, dtype=np.uint8)]
    pool = (nib[0::2] | (nib[1::2] << 4)).astype(np.uint8)
    return PulseBatch(kind, rec, pool, rssi, clock, domain)


def unpack_pulse(batch: PulseBatch, i: int) -> Dict[str, Any]:
    """Inverse of :func:`pack_pulse` for one message (used by corpus tools and tests).

    Returns a parser-style dict that packs back to the same record.  SP is synthesised
    (the reference validates but never uses it, message_synced.py:35-39).
    """
    r = batch.msgs[i]
    d: Dict[str, Any] = {}
    if not (r["flags"] & MSG_VALID):
        return {"data": ""}
    for s in range(int(r["npat"])):
        d[f"P{(int(r['pat_ids']) >> (4 * s)) & 0xF}"] = str(int(r["pat"][s]))
    base = int(r["doff"]) * 16
    dl = int(r["dlen"])
    by = batch.digits[base : base + (dl + 1) // 2]
    nib = np.empty(len(by) * 2, dtype=np.uint8)
    nib[0::2] = by & 0xF
    nib[1::2] = by >> 4
    d["data"] = "".join("0123456789??????"[v] for v in nib[:dl])
    if batch.kind == KIND_MS:
        cp = int(r["cp"])
        d["CP"] = str((int(r["pat_ids"]) >> (4 * cp)) & 0xF) if cp != 0xFF else "9"
        d["SP"] = "0"
    if batch.rssi[i] is not None:
        d["R"] = batch.rssi[i]
    return d


class HexBatch:
    """Packed MC or MN batch."""

    __slots__ = ("kind", "msgs", "digits", "n", "protocol_ids", "data", "domain")

    def __init__(self, kind, msgs, digits, protocol_ids, data, domain=None):
        self.kind = kind
        self.msgs = msgs
        self.digits = digits
        self.protocol_ids = protocol_ids
        self.data = data
        self.n = len(msgs)
        self.domain = domain or {}   # message index -> why it is outside the packed domain

    def algorithmic_input_bytes(self) -> int:
        return int(16 * self.n + ((self.msgs["hlen"].astype(np.int64) + 1) // 2).sum())


def pack_hex(msgs: Sequence[Dict[str, Any]], kind: int, proto_index: Dict[str, int],
             toggle_polarity: bool = False, method_override: int = 0, strict: bool = False) -> HexBatch:
    """Pack MC (kind 2) or MN (kind 3) dicts: protocol_id, data (hex), clock, bit_length.

    D must be upper-case hex (what the firmware emits): the reference's polarity inversion is an
    upper-case-only ``str.translate`` (manchester.py:36) and several MN converters echo the input
    verbatim, neither of which a 4-bit nibble can express -> such a message is outside the packed
    domain (SDB_MSG_DOMAIN, status "DomainError" for that message; ``strict`` raises instead).
    """
    n = len(msgs)
    rec = np.zeros(n, dtype=HEX_DTYPE)
    rec["proto"] = 0xFFFF
    streams: List[np.ndarray] = []
    pids: List[Any] = []
    datas: List[Any] = []
    domain: Dict[int, str] = {}
    empty = np.zeros(0, dtype=np.uint8)
    for i, m in enumerate(msgs):
        pid = m.get("protocol_id")
        pids.append(pid)
        data = m.get("data", "") if kind == KIND_MC else m.get("data")
        datas.append(data)
        ok = pid in proto_index if isinstance(pid, str) else False
        if kind == KIND_MN and "protocol_id" not in m:
            ok = False
        if method_override:                            # direct Conv*(msg_data): the converter never looks at the table
            ok = True
        if not ok:
            streams.append(empty)
            continue
        rec["proto"][i] = proto_index[pid] if not method_override else 0
        rec["rsv"][i] = method_override
        if data is None:
            data = ""
        why = None
        nib = empty
        clock = bitlen = 0
        if not isinstance(data, str):
            why = "'data' must be a str"
        else:
            raw = data.encode("ascii", "replace")
            b = np.frombuffer(raw, dtype=np.uint8)
            nib = _HEX_LUT[b]
            if len(raw) > MAX_HEX:
                why = f"D has {len(raw)} hex characters (max {MAX_HEX})"
            elif (nib == 0xFF).any():
                why = "D is not a hex string"
            elif ((b >= ord("a")) & (b <= ord("f"))).any():
                why = "D must be upper-case hex (what the firmware emits)"
        if why is None and kind == KIND_MC:
            clock = m.get("clock", 0)
            bitlen = m.get("bit_length", 0)
            if not isinstance(clock, int) or not isinstance(bitlen, int):
                why = "clock / bit_length must be int"
        if why is not None:
            if strict:
                raise DomainError(f"message {i}: {why}")
            domain[i] = why
            rec["flags"][i] = MSG_DOMAIN
            streams.append(empty)
            continue
        rec["flags"][i] = MSG_VALID | (HEX_TOGGLE_POLARITY if toggle_polarity else 0)
        streams.append(nib)
        rec["hlen"][i] = len(nib)
        if kind == KIND_MC:
            rec["clock"][i] = max(-(2**31), min(2**31 - 1, clock))
            rec["bitlen"][i] = max(-32768, min(32767, bitlen))
    pool, doff = pack_digit_streams(streams)
    rec["doff"] = doff
    return HexBatch(kind, rec, pool, pids, datas, domain)


def unpack_hex(batch: HexBatch, i: int) -> Dict[str, Any]:
    """Inverse of :func:`pack_hex` for one message (corpus tools and tests)."""
    r = batch.msgs[i]
    if not (r["flags"] & MSG_VALID):
        return {"data": ""}
    base = int(r["doff"]) * 16
    hl = int(r["hlen"])
    by = batch.digits[base : base + (hl + 1) // 2]
    nib = np.empty(len(by) * 2, dtype=np.uint8)
    nib[0::2] = by & 0xF
    nib[1::2] = by >> 4
    d: Dict[str, Any] = {"protocol_id": batch.protocol_ids[i], "data": "".join("0123456789ABCDEF"[v] for v in nib[:hl])}
    if batch.kind == KIND_MC:
        d["clock"] = int(r["clock"])
        d["bit_length"] = int(r["bitlen"])
    return d
