"""Protocol-table compiler: protocol dict -> flat device blob for libsdb200.so.

Replaces the per-message property lookups of the reference engine
(``check_property`` / ``get_property``, sd_protocols/sd_protocols.py:54-58) by a table that
is compiled ONCE per protocol set and copied to the GPU:

* every template value of ``sync/start/one/zero/float`` becomes an integer **tenths interval**
  ``[lo, hi]`` of accepted normalised pulse values plus a **gap-rank table**.  Both are produced
  by executing the reference's own float64 expressions (pattern_utils.py:15-26, :74-76, :83) in
  Python here, so the device needs integer compares only and stays bit-exact by construction
  (17 of 81 intervals differ from naive decimal arithmetic).
* ``modulematch`` regexes (message_unsynced.py:277-280) become fixed-offset character-class
  programs; the part that only re-checks the protocol's own preamble is folded away here.
* MS never-hit entries (string ``sync`` of the MN protocols, message_synced.py:114-118) are dropped.

Blob layout = ``SdbTblHeader`` + sections, mirrored by pysignalduino_b200/csrc/sdb_table.h.
"""
from __future__ import annotations

import re
from typing import Any, Dict, List, Optional, Tuple

import numpy as np

try:  # Python >= 3.11
    import re._parser as sre_parse  # type: ignore
    import re._constants as sre_c  # type: ignore
except ImportError:  # pragma: no cover
    import sre_parse  # type: ignore
    import sre_constants as sre_c  # type: ignore

TBL_MAGIC = 0x31424453  # "SDB1"
TBL_VERSION = 10

MAX_UNIQ = 4
MAX_TPL = 14
MAX_REQ = 12
KILL_WORDS = 8        # 32-byte kill-mask rows: up to 256 protocols per class

# flags of SdbPulseProto
PF_RECONSTRUCT = 0x01
PF_DISPATCH_BIN = 0x02
PF_REMOVE_ZERO = 0x04
PF_MM_END = 0x08      # modulematch ends with '$'
PF_MM_NEVER = 0x10    # modulematch can never match this protocol's preamble
PF_HAS_LIR_MAX = 0x20
PF_MM_HOST = 0x40

PD_IDS = {
    "postDemo_EM": 1, "postDemo_Revolt": 2, "postDemo_FS20": 3, "postDemo_FHT80": 4,
    "postDemo_FHT80TF": 5, "postDemo_WS2000": 6, "postDemo_WS7035": 7, "postDemo_WS7053": 8,
    "postDemo_lengtnPrefix": 9,
}
METHOD_IDS = {
    "mcBit2Funkbus": 1, "mcBit2Sainlogic": 2, "mcBit2AS": 3, "mcBit2Hideki": 4, "mcBit2Maverick": 5,
    "mcBit2OSV1": 6, "mcBit2OSV2o3": 7, "mcBit2OSPIR": 8, "mcRaw": 9, "mcraw": 10, "mcBit2TFA": 11,
    "mcBit2Grothe": 12, "mcBit2SomfyRTS": 13,
    "ConvBresser_lightning": 14, "ConvBresser_5in1": 15, "ConvBresser_6in1": 16, "ConvBresser_7in1": 17,
    "ConvPCA301": 18, "ConvKoppFreeControl": 19, "ConvLaCrosse": 20,
}
M_UNKNOWN = 21

KEYTPL_DTYPE = np.dtype(
    [
        ("len", "u1"), ("nuniq", "u1"), ("rsv", "<u2"),
        ("uidx", "<u4"),
        ("lo", "<i2", (MAX_UNIQ,)), ("hi", "<i2", (MAX_UNIQ,)),
        ("rank_off", "<u4", (MAX_UNIQ,)),
        ("vidx", "<u2", (MAX_UNIQ,)),          # MU: slot of (clock, interval) in the per-message candidate-mask table
    ]
)
assert KEYTPL_DTYPE.itemsize == 48

PULSEPROTO_DTYPE = np.dtype(
    [
        ("key", KEYTPL_DTYPE, (4,)),
        ("clock", "<f8"),
        ("proto", "<u2"), ("clk_idx", "<u2"),
        ("regex_min", "<i2"), ("lir_min", "<i2"), ("lir_max", "<i2"), ("mu_len_max", "<i2"),
        ("width", "u1"), ("padbits", "u1"), ("postdemod", "u1"), ("flags", "u1"),
        ("pre_len", "u1"), ("post_len", "u1"), ("mm_off", "<u2"),
        ("preamble", "S16"), ("postamble", "S4"),
        ("mm_nitems", "u1"), ("rsv", "u1", (7,)),
    ]
)
assert PULSEPROTO_DTYPE.itemsize == 248, PULSEPROTO_DTYPE.itemsize

PREFILTER_DTYPE = np.dtype([("clk_idx", "<u2"), ("nreq", "<u2"), ("vreq", "<u2", (MAX_REQ,))])
assert PREFILTER_DTYPE.itemsize == 28

VALROW_DTYPE = np.dtype([("clk_idx", "<u2"), ("lo", "<i2"), ("hi", "<i2"), ("rsv", "<u2")])
assert VALROW_DTYPE.itemsize == 8
MAX_VALS = 512

MMITEM_DTYPE = np.dtype([("mask", "<u4", (4,)), ("min", "<u2"), ("max", "<u2")])
assert MMITEM_DTYPE.itemsize == 20

HEXPROTO_DTYPE = np.dtype(
    [
        ("length_min", "<i4"), ("length_max", "<i4"),
        ("clock_min", "<i4"), ("clock_max", "<i4"),
        ("method", "u1"), ("flags", "u1"), ("pre_len", "u1"), ("pid_int", "u1"),
        ("preamble", "S16"),
    ]
)
assert HEXPROTO_DTYPE.itemsize == 36
HF_EXISTS = 0x01
HF_HAS_MIN = 0x02
HF_HAS_MAX = 0x04
HF_CLOCKRANGE = 0x08
HF_INVERT = 0x10
HF_MAX_IS_STR = 0x20
HF_IS_119 = 0x40

HEADER_DTYPE = np.dtype(
    [
        ("magic", "<u4"), ("version", "<u4"), ("nproto", "<u4"),
        ("n_ms", "<u4"), ("n_mu", "<u4"), ("n_clk", "<u4"), ("n_rank", "<u4"), ("n_mm", "<u4"),
        ("off_ms", "<u4"), ("off_mu", "<u4"), ("off_ms_pf", "<u4"), ("off_mu_pf", "<u4"),
        ("off_clk", "<u4"), ("off_rank", "<u4"), ("off_mm", "<u4"), ("off_hex", "<u4"),
        ("total", "<u4"), ("n_vals", "<u4"), ("off_vals", "<u4"), ("n_mu_vals", "<u4"),
        ("off_kill", "<u4"), ("rsv0", "<u4"), ("rsv1", "<u4"), ("rsv2", "<u4"),
    ]
)
assert HEADER_DTYPE.itemsize == 96


# --------------------------------------------------------------------------------------------
# float64 semantics of the reference, evaluated here once per template value
# --------------------------------------------------------------------------------------------
def calculate_tolerance(val: float) -> float:
    """pattern_utils.calculate_tolerance (pattern_utils.py:15-26), same float expressions."""
    abs_val = abs(val)
    if abs_val > 3:
        if abs_val > 16:
            return abs_val * 0.18
        else:
            return abs_val * 0.3
    return 1.0


def _accepts(t: int, s: float, tol: float) -> Tuple[bool, float]:
    """Would pattern_exists accept a slot whose normalised value is t/10 for template s?

    The normalised value is ``round(p / clock, 1)`` (message_synced.py:72, message_unsynced.py:64),
    i.e. the double nearest to the decimal t/10, which is what ``t / 10`` evaluates to.
    """
    gap = abs(t / 10 - s)                       # pattern_utils.py:74
    return (gap <= 0.001 or gap <= tol), gap    # pattern_utils.py:75


def tenths_interval(s: float) -> Tuple[int, int, List[int]]:
    """Accepted tenths interval [lo, hi] of template value s and the dense gap rank of every t in it."""
    tol = calculate_tolerance(s)
    centre = int(round(s * 10))
    span = int(tol * 10) + 30
    acc = [t for t in range(centre - span, centre + span + 1) if _accepts(t, s, tol)[0]]
    if not acc:
        raise NotImplementedError(f"template value {s} accepts no tenths value")
    lo, hi = acc[0], acc[-1]
    if acc != list(range(lo, hi + 1)):
        raise NotImplementedError(f"accept set of template value {s} is not contiguous")
    if _accepts(lo - 1, s, tol)[0] or _accepts(hi + 1, s, tol)[0] or not (-32000 < lo and hi < 32000):
        raise NotImplementedError(f"accept interval of template value {s} out of range")
    gaps = [_accepts(t, s, tol)[1] for t in range(lo, hi + 1)]
    order = sorted(set(gaps))
    rank_of = {g: i for i, g in enumerate(order)}     # equal float gaps share a rank -> stable tie-break by slot
    return lo, hi, [rank_of[g] for g in gaps]


# --------------------------------------------------------------------------------------------
# modulematch compiler
# --------------------------------------------------------------------------------------------
def _charset_mask(node) -> int:
    """128-bit ASCII membership mask of one regex atom (LITERAL / ANY / IN)."""
    op, av = node
    if op == sre_c.LITERAL:
        if av >= 128:
            raise NotImplementedError("non-ASCII literal in modulematch")
        return 1 << av
    if op == sre_c.ANY:
        return ((1 << 128) - 1) & ~(1 << 10)          # '.' excludes newline
    if op == sre_c.IN:
        mask, negate = 0, False
        for iop, iav in av:
            if iop == sre_c.NEGATE:
                negate = True
            elif iop == sre_c.LITERAL:
                mask |= 1 << iav
            elif iop == sre_c.RANGE:
                for c in range(iav[0], iav[1] + 1):
                    mask |= 1 << c
            else:
                raise NotImplementedError(f"unsupported class item {iop} in modulematch")
        if negate:
            mask = ((1 << 128) - 1) & ~mask
        return mask
    raise NotImplementedError(f"unsupported regex atom {op} in modulematch")


def compile_modulematch(pattern: str) -> Tuple[List[Tuple[int, int, int]], bool]:
    """``^`` + atoms with fixed counts + at most one trailing variable atom + optional ``$``.

    Returns ([(mask128, min, max)], end_anchored).  max 0xFFFF = unbounded.
    """
    tree = list(sre_parse.parse(pattern))
    if not tree or tree[0] != (sre_c.AT, sre_c.AT_BEGINNING):
        raise NotImplementedError(f"modulematch {pattern!r} is not ^-anchored")
    tree = tree[1:]
    end = False
    if tree and tree[-1] == (sre_c.AT, sre_c.AT_END):
        end = True
        tree = tree[:-1]
    items: List[Tuple[int, int, int]] = []
    for node in tree:
        op, av = node
        if op in (sre_c.MAX_REPEAT, sre_c.MIN_REPEAT):
            lo, hi, sub = av
            sub = list(sub)
            if len(sub) != 1:
                raise NotImplementedError(f"modulematch {pattern!r}: repeated group")
            mask = _charset_mask(sub[0])
            hi = 0xFFFF if hi == sre_c.MAXREPEAT else hi
            items.append((mask, lo, hi))
        elif op in (sre_c.LITERAL, sre_c.ANY, sre_c.IN):
            items.append((_charset_mask(node), 1, 1))
        else:
            raise NotImplementedError(f"modulematch {pattern!r}: unsupported construct {op}")
    for mask, lo, hi in items[:-1]:
        if lo != hi:
            raise NotImplementedError(f"modulematch {pattern!r}: variable repeat before the last atom")
    return items, end


def fold_preamble(items, end: bool, preamble: str):
    """Evaluate the part of the program that only looks at the (constant) preamble.

    Returns (remaining_items, never) where remaining items apply from payload offset len(preamble).
    Folding stops at the first variable item.
    """
    pos = 0
    out = list(items)
    while out and pos < len(preamble):
        mask, lo, hi = out[0]
        if lo != hi:
            break
        take = min(lo, len(preamble) - pos)
        for k in range(take):
            c = ord(preamble[pos + k])
            if c >= 128 or not (mask >> c) & 1:
                return [], True
        pos += take
        if take == lo:
            out.pop(0)
        else:
            out[0] = (mask, lo - take, lo - take)
    if pos < len(preamble):
        # a variable item (or the end of the program) starts inside the preamble: keep the device
        # program simple by refusing shapes that never occur in the shipped table
        if out:
            mask, lo, hi = out[0]
            rest = preamble[pos:]
            ok = all(ord(c) < 128 and (mask >> ord(c)) & 1 for c in rest)
            if hi == 0xFFFF and len(out) == 1 and not end:
                # "X*" / ".*" tail: matches any number of X then stops; anything may follow -> only min matters
                n_ok = 0
                for c in rest:
                    if ord(c) < 128 and (mask >> ord(c)) & 1:
                        n_ok += 1
                    else:
                        break
                if n_ok >= lo:
                    return [], False
                if n_ok < len(rest):
                    return [], True
                return [(mask, lo - n_ok, 0xFFFF)], False
            raise NotImplementedError("modulematch variable atom overlapping the preamble")
        if end:
            return [], True          # '$' inside the preamble
        return [], False
    return out, False


# --------------------------------------------------------------------------------------------
# table compiler
# --------------------------------------------------------------------------------------------
class CompiledTable:
    def __init__(self, blob: bytes, ids: List[str], ms_ids: List[str], mu_ids: List[str], info: Dict[str, Any],
                 hex_rows: Optional[np.ndarray] = None, unsupported: Optional[Dict[str, str]] = None):
        self.blob = blob
        self.ids = ids              # protocol ids in table order (hit.proto indexes this)
        self.ms_ids = ms_ids
        self.mu_ids = mu_ids
        self.info = info
        self.hex_rows = hex_rows    # HEXPROTO_DTYPE row per protocol id (length rules, method, preamble)
        # "<id> (MS|MU|MC/MN)" -> why that row is NOT in the device table (the protocol never matches in that class).
        # Empty for the shipped table; a user-edited table degrades per protocol instead of failing engine creation.
        self.unsupported = unsupported or {}

    def pulse_rows(self, kind: str) -> np.ndarray:
        """The compiled MS / MU protocol rows (PULSEPROTO_DTYPE) as a read-only view of the blob."""
        hdr = np.frombuffer(self.blob, dtype=HEADER_DTYPE, count=1)[0]
        n, off = (int(hdr["n_ms"]), int(hdr["off_ms"])) if kind == "MS" else (int(hdr["n_mu"]), int(hdr["off_mu"]))
        return np.frombuffer(self.blob, dtype=PULSEPROTO_DTYPE, count=n, offset=off)


class _RankPool:
    def __init__(self):
        self.data: List[int] = []
        self.cache: Dict[float, Tuple[int, int, int]] = {}

    def get(self, s: float) -> Tuple[int, int, int]:
        if s not in self.cache:
            lo, hi, ranks = tenths_interval(s)
            off = len(self.data)
            self.data.extend(ranks)
            self.cache[s] = (lo, hi, off)
        return self.cache[s]


def _fill_key(rec, values, pool: _RankPool, what: str):
    vals = [float(x) for x in values]
    if len(vals) > MAX_TPL:
        raise NotImplementedError(f"{what}: template longer than {MAX_TPL}")
    uniq: List[float] = []
    for v in vals:
        if v not in uniq:                      # pattern_utils.py:54-57 (== on floats)
            uniq.append(v)
    if len(uniq) > MAX_UNIQ:
        raise NotImplementedError(f"{what}: more than {MAX_UNIQ} distinct template values")
    rec["len"] = len(vals)
    rec["nuniq"] = len(uniq)
    uidx = 0
    for i, v in enumerate(vals):
        uidx |= uniq.index(v) << (2 * i)
    rec["uidx"] = uidx
    for u, s in enumerate(uniq):
        lo, hi, off = pool.get(s)
        if lo < -15000 or hi > 15000:          # the kernels keep tenths as biased 15-bit values (csrc/sdb_pulse.cu: T_BIAS / T_CLAMP)
            raise NotImplementedError(f"{what}: accept interval [{lo}, {hi}] tenths outside +-15000")
        rec["lo"][u], rec["hi"][u], rec["rank_off"][u] = lo, hi, off


def _as_floats(v) -> Optional[List[float]]:
    try:
        return [float(x) for x in v]
    except (TypeError, ValueError):
        return None


def compile_table(protocols: Dict[str, Dict[str, Any]], strict: bool = False) -> CompiledTable:
    """Compile a protocol dict (``SDProtocols._protocols``) into the device blob.

    Tolerant like the reference (which skips what it cannot use, message_synced.py:206, message_unsynced.py:234): a row
    whose shape the device layout cannot hold (more than 4 distinct template values, symbol widths other than 1 / 2 / 4, ...)
    is left out and listed in ``CompiledTable.unsupported``; a ``modulematch`` outside the compiled program's shapes is
    evaluated by the host formatter (PF_MM_HOST).  ``strict=True`` raises ``NotImplementedError`` instead.
    """
    ids = list(protocols.keys())
    if len(ids) >= 0xFFFF:
        raise NotImplementedError("too many protocols")
    pool = _RankPool()
    mm_items: List[Tuple[int, int, int]] = []
    unsupported: Dict[str, str] = {}

    def guarded(label: str, fn):
        """Run one row compiler; an unsupported shape drops that row only."""
        if strict:
            return fn()
        try:
            return fn()
        except (NotImplementedError, ValueError, TypeError, OverflowError) as e:
            unsupported[label] = f"{type(e).__name__}: {e}"
            return None

    # ---------------- MS: get_keys('sync') (message_synced.py:79) ----------------
    ms_rows, ms_pf, ms_ids = [], [], []

    def ms_row(idx, pid, pr):
        sync = pr.get("sync")
        if not sync:
            raise NotImplementedError(f"protocol {pid}: falsy 'sync'")
        svals = _as_floats(sync)
        if svals is None:
            return None                                # float('D') -> match_failed (:114-118): never hits
        one = pr.get("one")
        if not one:
            # signal_width 0 (:106-107): fails `length_min > 0` (:150-156) unless length_min <= 0,
            # in which case range(..., 0) raises (:174)
            if int(pr.get("length_min", -1)) > 0:
                return None
            raise NotImplementedError(f"protocol {pid}: MS protocol without 'one' and length_min <= 0")
        rec = np.zeros((), dtype=PULSEPROTO_DTYPE)
        _fill_common(rec, idx, pid, pr, pool, ms=True)
        _fill_key(rec["key"][0], svals, pool, f"{pid}.sync")
        rec["clock"] = float(pr.get("clockabs", 0))    # :83
        rec["regex_min"] = int(pr.get("length_min", -1))   # :152
        if pr.get("postDemodulation") and rec["postdemod"] and pr.get("float"):
            raise NotImplementedError(f"protocol {pid}: MS postDemodulation with 'float' symbols raises in the reference")
        return rec

    for idx, (pid, pr) in enumerate(protocols.items()):
        if "sync" not in pr:
            continue
        rec = guarded(f"{pid} (MS)", lambda: ms_row(idx, pid, pr))
        if rec is not None:
            ms_rows.append(rec)
            ms_ids.append(pid)

    # ---------------- MU: get_keys('clockabs') + active (message_unsynced.py:45-49) ----------------
    mu_rows, mu_pf, mu_ids, clocks = [], [], [], []
    mu_vals: Dict[Tuple[int, int, int], int] = {}

    def mu_row(idx, pid, pr):
        rec = np.zeros((), dtype=PULSEPROTO_DTYPE)
        _fill_common(rec, idx, pid, pr, pool, ms=False)
        start = pr.get("start")
        if start and isinstance(start, list):          # :71
            _fill_key(rec["key"][0], start, pool, f"{pid}.start")
        clock = float(pr.get("clockabs", 1))           # :59
        if clock == 0:
            raise NotImplementedError(f"protocol {pid}: clockabs 0 (ZeroDivisionError in the reference)")
        rec["clock"] = clock
        lmin = pr.get("length_min", 0)                 # :178 goes into the regex text {MIN,}
        if not re.fullmatch(r"\d+", str(lmin)):
            raise NotImplementedError(f"protocol {pid}: length_min {lmin!r} is not a regex repeat count")
        rec["regex_min"] = int(str(lmin))
        if int(rec["regex_min"]) == 0 and int(rec["key"][0]["len"]) == 0:
            raise NotImplementedError(f"protocol {pid}: empty-match regex (no start, length_min 0)")
        lmax = pr.get("length_max", None)              # :197,:217 truthiness
        rec["mu_len_max"] = int(lmax) if lmax else -1
        new_items: List[Tuple[int, int, int]] = []
        mm = pr.get("modulematch")
        if mm:                                         # :277-280
            try:
                items, end = compile_modulematch(mm)
                items, never = fold_preamble(items, end, str(pr.get("preamble", "")))
            except (NotImplementedError, re.error, ValueError):
                if strict:
                    raise
                re.compile(mm)                         # a regex Python itself rejects raises in the reference too
                rec["flags"] |= PF_MM_HOST             # hits are flagged; the host formatter runs re.search
                return rec, new_items
            if not end and items:
                # re.search without '$': a trailing {lo,hi} only needs its minimum; {0,..} is a no-op
                mask, lo, _hi = items[-1]
                items[-1] = (mask, lo, lo)
                if lo == 0:
                    items.pop()
            merged: List[Tuple[int, int, int]] = []
            for it in items:                           # "......" -> one atom with count 6 (fixed-count atoms only)
                if merged and merged[-1][0] == it[0] and merged[-1][1] == merged[-1][2] and it[1] == it[2]:
                    merged[-1] = (it[0], merged[-1][1] + it[1], merged[-1][2] + it[2])
                else:
                    merged.append(it)
            items = merged
            if never:
                rec["flags"] |= PF_MM_NEVER
            elif items or end:
                if len(items) > 255:
                    raise NotImplementedError(f"protocol {pid}: modulematch program too long")
                rec["mm_nitems"] = len(items)
                new_items = list(items)
                if end:
                    rec["flags"] |= PF_MM_END
                if not items:
                    rec["mm_off"] = 0xFFFE             # '$' only: patched to a valid offset below
        return rec, new_items

    for idx, (pid, pr) in enumerate(protocols.items()):
        if "clockabs" not in pr:
            continue
        if not pr.get("active", True):
            continue
        if not pr.get("one"):
            continue                                   # signal_width == 0 -> every match skipped (:205)
        got = guarded(f"{pid} (MU)", lambda: mu_row(idx, pid, pr))
        if got is None:
            continue
        rec, new_items = got
        clock = float(rec["clock"])
        if clock not in clocks:
            clocks.append(clock)
        rec["clk_idx"] = clocks.index(clock)
        if new_items or int(rec["mm_off"]) == 0xFFFE:
            rec["mm_off"] = len(mm_items)
            mm_items.extend(new_items)
        # every (clock, accept interval) pair gets one slot of the per-message candidate-mask table
        for kk in range(4):
            kt = rec["key"][kk]
            for u in range(int(kt["nuniq"])):
                vkey = (int(rec["clk_idx"]), int(kt["lo"][u]), int(kt["hi"][u]))
                if vkey not in mu_vals:
                    mu_vals[vkey] = len(mu_vals)
                kt["vidx"][u] = mu_vals[vkey]
        mu_rows.append(rec)
        mu_ids.append(pid)

    # MS rows use the same per-message mask table: their intervals follow the MU pairs, with clock slot 0
    # (the MS kernel puts the message's own tenths, normalised by P[CP], into T[0])
    ms_vals: Dict[Tuple[int, int], int] = {}
    for rec in ms_rows:
        for kk in range(4):
            kt = rec["key"][kk]
            for u in range(int(kt["nuniq"])):
                vkey2 = (int(kt["lo"][u]), int(kt["hi"][u]))
                if vkey2 not in ms_vals:
                    ms_vals[vkey2] = len(ms_vals)
                kt["vidx"][u] = len(mu_vals) + ms_vals[vkey2]

    # ---------------- MC / MN protocol rows (every protocol id, table order) ----------------
    hexrows = np.zeros(len(ids), dtype=HEXPROTO_DTYPE)

    def hex_row(idx, pid, pr):
        h = np.zeros((), dtype=HEXPROTO_DTYPE)
        fl = HF_EXISTS
        if pr.get("length_min") is not None:
            fl |= HF_HAS_MIN
            h["length_min"] = int(pr["length_min"])
        if pr.get("length_max") is not None:
            fl |= HF_HAS_MAX
            h["length_max"] = int(pr["length_max"])
            if isinstance(pr["length_max"], str):
                fl |= HF_MAX_IS_STR
        cr = pr.get("clockrange")
        if cr and len(cr) >= 2:
            fl |= HF_CLOCKRANGE
            h["clock_min"], h["clock_max"] = int(cr[0]), int(cr[1])
        if pr.get("polarity", "") == "invert":
            fl |= HF_INVERT
        if pid == "119":
            fl |= HF_IS_119
        meth = pr.get("method")
        if meth:
            h["method"] = METHOD_IDS.get(meth.split(".")[-1], M_UNKNOWN)
        h["flags"] = fl
        try:
            h["pid_int"] = 1 if int(pid) > 0 else 2    # as shipped, mcRaw receives the id as its mcbitnum (manchester.py:120)
        except ValueError:
            h["pid_int"] = 0
        pre = str(pr.get("preamble", "")).encode("latin-1")
        if len(pre) > 16:
            raise NotImplementedError(f"protocol {pid}: preamble too long")
        h["preamble"], h["pre_len"] = pre, len(pre)
        return h

    for idx, (pid, pr) in enumerate(protocols.items()):
        h = guarded(f"{pid} (MC/MN)", lambda: hex_row(idx, pid, pr))
        if h is not None:
            hexrows[idx] = h                           # a dropped row keeps flags == 0: "protocol does not exist" on the device

    # ---------------- assemble ----------------
    def arr(rows, dtype):
        a = np.zeros(len(rows), dtype=dtype)
        for i, r in enumerate(rows):
            a[i] = r
        return a

    ms_arr, mu_arr = arr(ms_rows, PULSEPROTO_DTYPE), arr(mu_rows, PULSEPROTO_DTYPE)
    ms_pf = [_prefilter(rec, keys=(0, 1, 2)) for rec in ms_rows]
    mu_pf = [_prefilter(rec, keys=(0, 1, 2)) for rec in mu_rows]
    ms_pf_arr, mu_pf_arr = arr(ms_pf, PREFILTER_DTYPE), arr(mu_pf, PREFILTER_DTYPE)
    # n_clk clocks followed by n_clk values of 10/clock (device fast path of round(p/clock, 1), exact path on near-ties)
    clk_arr = np.asarray(clocks + [10.0 / c for c in clocks], dtype="<f8")
    rank_arr = np.asarray(pool.data, dtype="<u2")
    mm_arr = np.zeros(len(mm_items), dtype=MMITEM_DTYPE)
    for i, (mask, lo, hi) in enumerate(mm_items):
        for w in range(4):
            mm_arr["mask"][i, w] = (mask >> (32 * w)) & 0xFFFFFFFF
        mm_arr["min"][i], mm_arr["max"][i] = lo, hi

    hdr = np.zeros((), dtype=HEADER_DTYPE)
    sections = []
    off = HEADER_DTYPE.itemsize

    def add(a) -> int:
        nonlocal off
        off = (off + 15) // 16 * 16
        start = off
        sections.append((start, a.tobytes()))
        off += a.nbytes
        return start

    hdr["magic"], hdr["version"], hdr["nproto"] = TBL_MAGIC, TBL_VERSION, len(ids)
    hdr["n_ms"], hdr["n_mu"], hdr["n_clk"] = len(ms_arr), len(mu_arr), len(clocks)
    hdr["n_rank"], hdr["n_mm"] = len(rank_arr), len(mm_arr)
    hdr["off_ms"] = add(ms_arr)
    hdr["off_mu"] = add(mu_arr)
    hdr["off_ms_pf"] = add(ms_pf_arr)
    hdr["off_mu_pf"] = add(mu_pf_arr)
    hdr["off_clk"] = add(clk_arr)
    hdr["off_rank"] = add(rank_arr)
    hdr["off_mm"] = add(mm_arr)
    hdr["off_hex"] = add(hexrows)
    if len(mu_vals) + len(ms_vals) > MAX_VALS:
        raise NotImplementedError(f"{len(mu_vals) + len(ms_vals)} distinct (clock, interval) pairs (max {MAX_VALS})")
    val_arr = np.zeros(len(mu_vals) + len(ms_vals), dtype=VALROW_DTYPE)
    for (ck, lo, hi), i in mu_vals.items():
        val_arr[i] = (ck, lo, hi, 0)
    for (lo, hi), i in ms_vals.items():
        val_arr[len(mu_vals) + i] = (0, lo, hi, 0)
    hdr["n_vals"] = len(val_arr)
    hdr["n_mu_vals"] = len(mu_vals)
    hdr["off_vals"] = add(val_arr)
    # kill masks: bit r of row v = protocol row r (of its class) needs pair v to have a candidate slot
    # (the transposed prefilter: one lane per pair ORs the rows of the pairs without a candidate)
    kill = np.zeros((len(val_arr), KILL_WORDS), dtype="<u4")
    for pf_arr in (mu_pf_arr, ms_pf_arr):
        if len(pf_arr) > 32 * KILL_WORDS:
            raise NotImplementedError(f"{len(pf_arr)} protocols in one class (max {32 * KILL_WORDS})")
        for r, pf in enumerate(pf_arr):
            for i in range(int(pf["nreq"])):
                kill[int(pf["vreq"][i]), r >> 5] |= np.uint32(1 << (r & 31))
    hdr["off_kill"] = add(kill)
    total = (off + 15) // 16 * 16
    hdr["total"] = total
    blob = bytearray(total)
    blob[: HEADER_DTYPE.itemsize] = hdr.tobytes()
    for start, b in sections:
        blob[start : start + len(b)] = b
    info = {"n_ms": len(ms_arr), "n_mu": len(mu_arr), "n_clk": len(clocks), "n_rank": len(rank_arr), "n_vals": len(mu_vals) + len(ms_vals), "n_mu_vals": len(mu_vals),
            "n_mm_items": len(mm_arr), "bytes": total, "clocks": clocks}
    return CompiledTable(bytes(blob), ids, ms_ids, mu_ids, info, hexrows, unsupported)


def _fill_common(rec, idx: int, pid: str, pr: Dict[str, Any], pool: _RankPool, ms: bool) -> None:
    rec["proto"] = idx
    rec["mm_off"] = 0xFFFF
    for k, name in ((1, "one"), (2, "zero"), (3, "float")):
        v = pr.get(name)
        if not v:
            continue                                   # `if not search_pattern: continue`
        fv = _as_floats(v)
        if fv is None:
            raise NotImplementedError(f"protocol {pid}: non-numeric '{name}'")
        _fill_key(rec["key"][k], fv, pool, f"{pid}.{name}")
    w = len(pr["one"])
    for k in (2, 3):
        if int(rec["key"][k]["len"]) not in (0, w):
            raise NotImplementedError(f"protocol {pid}: symbol widths differ (SURVEY App. B says they never do)")
    if w not in (1, 2, 4):
        raise NotImplementedError(f"protocol {pid}: symbol width {w}")
    rec["width"] = w
    lmin = pr.get("length_min", -1)                    # helpers.length_in_range (helpers.py:144-154)
    rec["lir_min"] = int(lmin) if lmin is not None else -1
    lmax = pr.get("length_max")                        # helpers.py:157-164
    if lmax is not None:
        rec["lir_max"] = int(lmax)
        rec["flags"] |= PF_HAS_LIR_MAX
    rec["mu_len_max"] = -1
    pad = int(pr.get("paddingbits", 4))
    if pad <= 0 or pad > 64:
        raise NotImplementedError(f"protocol {pid}: paddingbits {pad}")
    rec["padbits"] = pad
    pd = pr.get("postDemodulation")
    if pd:
        rec["postdemod"] = PD_IDS.get(pd.split(".")[-1], 0)     # hasattr gate: unknown names are skipped
    if pr.get("reconstructBit"):
        rec["flags"] |= PF_RECONSTRUCT
    if int(pr.get("dispatchBin", 0)) == 1:
        rec["flags"] |= PF_DISPATCH_BIN
    if pr.get("remove_zero", 0):
        rec["flags"] |= PF_REMOVE_ZERO
        if pr.get("float"):
            raise NotImplementedError(f"protocol {pid}: remove_zero with 'float' raises AttributeError in the reference")
    pre = str(pr.get("preamble", "")).encode("latin-1")
    post = str(pr.get("postamble", "")).encode("latin-1")
    if len(pre) > 16 or len(post) > 4:
        raise NotImplementedError(f"protocol {pid}: preamble/postamble too long")
    rec["preamble"], rec["pre_len"] = pre, len(pre)
    rec["postamble"], rec["post_len"] = post, len(post)


def _prefilter(rec, keys) -> np.ndarray:
    """Every unique value of the mandatory keys needs >= 1 candidate slot (pattern_utils.py:78-80):
    the rows of the per-message candidate-mask table that must be non-zero for the protocol to stay alive."""
    pf = np.zeros((), dtype=PREFILTER_DTYPE)
    pf["clk_idx"] = rec["clk_idx"]
    req: List[int] = []
    for k in keys:
        kt = rec["key"][k]
        for u in range(int(kt["nuniq"])):
            v = int(kt["vidx"][u])
            if v not in req:
                req.append(v)
    if len(req) > MAX_REQ:
        raise NotImplementedError("prefilter overflow")
    for i, v in enumerate(req):
        pf["vreq"][i] = v
    pf["nreq"] = len(req)
    return pf
