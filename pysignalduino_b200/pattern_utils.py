"""``pattern_utils`` of the drop-in package (sd_protocols/pattern_utils.py:11-136).

``pattern_exists`` runs on the GPU: the call is compiled into one template row (accepted tenths intervals and gap ranks
derived by evaluating the reference's own float expressions, table.py) and resolved by the same warp-level code the MS /
MU kernels use (``resolve_key`` / ``resolve_general`` in csrc/sdb_pulse.cu) through ``sdb_unit_pattern_exists``.
Inputs the packed domain cannot hold raise ``DomainError`` — they are never answered differently from the reference:
  * more than 8 patterns, or an id that is not one digit '0'..'9';
  * a pattern value that is not a multiple of 0.1 (the demodulators only ever pass ``round(p / clock, 1)``) or beyond
    +-3200;
  * a search pattern of more than 14 pulses or more than 4 distinct values; data longer than 4096 characters.
"""
from __future__ import annotations

import itertools
from typing import Any, Dict, List, Union

import numpy as np

from . import pack
from .table import KEYTPL_DTYPE, MAX_TPL, MAX_UNIQ, calculate_tolerance, tenths_interval  # noqa: F401  (re-exported)

_handler = None


def _engine():
    global _handler
    if _handler is None:
        from .sd_protocols import SDProtocols

        _handler = SDProtocols()
    return _handler.engine()


def is_in_tolerance(val1: float, val2: float, tol: float) -> bool:
    """pattern_utils.py:11-13"""
    return abs(val1 - val2) <= tol


def cartesian_product(lists: List[List[Any]]) -> List[List[Any]]:
    """pattern_utils.py:28-32"""
    if not lists:
        return [[]]
    return [list(p) for p in itertools.product(*lists)]


def pattern_exists(search_pattern: List[float], pattern_list: Dict[str, float], raw_data: str, debug_callback=None) -> Union[str, int]:
    """pattern_utils.py:34-136: the id string of ``search_pattern`` if it occurs in ``raw_data``, else -1."""
    vals = [float(v) for v in search_pattern]
    if not vals:
        return ""                                    # the empty product yields the empty target, which every string contains
    uniq: List[float] = []
    for v in vals:
        if v not in uniq:
            uniq.append(v)
    if len(vals) > MAX_TPL or len(uniq) > MAX_UNIQ:
        raise pack.DomainError(f"search pattern with {len(vals)} pulses / {len(uniq)} distinct values (max {MAX_TPL} / {MAX_UNIQ})")
    if len(pattern_list) > pack.MAX_SLOTS:
        raise pack.DomainError(f"{len(pattern_list)} patterns (max {pack.MAX_SLOTS})")
    if not isinstance(raw_data, str) or len(raw_data) > pack.MAX_DIGITS:
        raise pack.DomainError("raw_data must be a str of at most 4096 characters")
    tenths = np.full(8, -32768, dtype=np.int16)
    ids = 0
    for s, (pid, pval) in enumerate(pattern_list.items()):
        if not (isinstance(pid, str) and len(pid) == 1 and pid in "0123456789"):
            raise pack.DomainError(f"pattern id {pid!r} is not a single digit")
        t = round(float(pval) * 10)
        if abs(t) > 32000 or t / 10 != float(pval):
            raise pack.DomainError(f"pattern value {pval!r} is not a multiple of 0.1 within +-3200")
        tenths[s] = t
        ids |= int(pid) << (4 * s)
    tpl = np.zeros((), dtype=KEYTPL_DTYPE)
    tpl["len"], tpl["nuniq"] = len(vals), len(uniq)
    uidx = 0
    for i, v in enumerate(vals):
        uidx |= uniq.index(v) << (2 * i)
    tpl["uidx"] = uidx
    rank: List[int] = []
    for u, sv in enumerate(uniq):
        try:
            lo, hi, ranks = tenths_interval(sv)
        except NotImplementedError as e:
            raise pack.DomainError(str(e)) from None
        tpl["lo"][u], tpl["hi"][u], tpl["rank_off"][u] = lo, hi, len(rank)
        rank.extend(ranks)
    raw = raw_data.encode("ascii", "replace")
    nib = pack._DIGIT_LUT[np.frombuffer(raw, dtype=np.uint8)] if raw else np.zeros(0, dtype=np.uint8)
    pool, _ = pack.pack_digit_streams([nib])
    found, digits, _pos = _engine().unit_pattern_exists(tpl, np.asarray(rank, dtype=np.uint16), tenths, ids, len(pattern_list),
                                                        pool, len(raw))
    if not found:
        return -1
    return "".join(str(d) for d in digits)
