"""Shared helpers of the test-suite."""
from __future__ import annotations

import gzip
import json
from pathlib import Path
from typing import Any, Dict, List, Tuple

GOLDEN = Path(__file__).resolve().parent / "golden"

Canonical = Tuple[str, List[Tuple[str, str, int]]]


def canonical_gpu(sdp, batch, res) -> List[Canonical]:
    """GPU result arrays -> [(status, [(protocol_id, payload, bit_length)...])] per message."""
    statuses, results = sdp.format_results(batch, res)
    out = []
    for st, lst in zip(statuses, results):
        out.append((st, [(str(r["protocol_id"]), r["payload"], int(r["meta"].get("bit_length", -1))) for r in lst]))
    return out


def load_golden(name: str) -> List[Dict[str, Any]]:
    with gzip.open(GOLDEN / name, "rt", encoding="utf-8") as f:
        return json.load(f)


def golden_expected(rec) -> Canonical:
    return (rec["status"], [(a, b, int(c)) for a, b, c in rec["results"]])


def diff_report(got: List[Canonical], exp: List[Canonical], limit: int = 5) -> str:
    bad = [i for i, (g, e) in enumerate(zip(got, exp)) if g != e]
    lines = [f"{len(bad)} of {len(exp)} messages differ"]
    for i in bad[:limit]:
        lines.append(f"  [{i}] got={got[i]}\n       exp={exp[i]}")
    return "\n".join(lines)


def compare_raw(sdp, batch, res, ora_status, ora_hits, ora_pool, check_bits=True) -> str:
    """Vectorised full comparison of a GPU result with the oracle's raw arrays; returns '' or a description.

    Device hits of one message are contiguous ([hit_off, hit_off+nhits)) and in reference order; messages are
    visited in batch order to line both sides up without Python loops over messages.
    """
    import numpy as np

    out, hits = res.out, res.hits
    if not np.array_equal(out["status"], ora_status):
        bad = np.nonzero(out["status"] != ora_status)[0]
        return f"status differs for {len(bad)} messages, first {bad[:5]}"
    nh = out["nhits"].astype(np.int64)
    starts = out["hit_off"].astype(np.int64)
    total = int(nh.sum())
    if total != len(hits):
        return f"sum(nhits)={total} != len(hits)={len(hits)}"
    # gather index: for message m the rows starts[m] .. starts[m]+nh[m]
    rep = np.repeat(np.arange(batch.n), nh)
    within = np.arange(total) - np.repeat(np.cumsum(nh) - nh, nh)
    order = np.repeat(starts, nh) + within
    g = hits[order]
    if not np.array_equal(g["msg"].astype(np.int64), rep):
        return "hit.msg does not match the per-message slots"
    keep = ~(((g["flags"] & 2) != 0) & (g["aux"] != 0))          # continuation rows of a TFA list carry no string
    pool, off = sdp.engine().format_hits(batch.kind, g, res.bits)
    off = off.astype(np.int64)
    glen = (off[1:] - off[:-1])[keep]
    gk = g[keep]
    if len(gk) != len(ora_hits):
        return f"{len(gk)} device hits vs {len(ora_hits)} oracle hits"
    if not np.array_equal(gk["msg"].astype(np.int64), ora_hits["msg"].astype(np.int64)):
        return "hit message indices differ"
    if not np.array_equal(gk["proto"].astype(np.int64), ora_hits["proto"].astype(np.int64)):
        bad = np.nonzero(gk["proto"].astype(np.int64) != ora_hits["proto"].astype(np.int64))[0]
        return f"protocol ids differ at {len(bad)} hits, first msg {int(gk['msg'][bad[0]])}"
    if check_bits and not np.array_equal(gk["nbits"].astype(np.int64), ora_hits["bit_length"].astype(np.int64)):
        return "bit_length differs"
    if not np.array_equal(glen, ora_hits["payload_len"].astype(np.int64)):
        bad = np.nonzero(glen != ora_hits["payload_len"].astype(np.int64))[0]
        return f"payload lengths differ at {len(bad)} hits, first msg {int(gk['msg'][bad[0]])}"
    if pool != ora_pool:
        return "payload bytes differ"
    return ""


def compare_payloads(batch, res, pool, ora_status, ora_hits, ora_pool) -> str:
    """The same comparison for a result of ``sdb_demod_host_payloads`` (12-byte SdbPayloadHit records + strings written by the
    DEVICE format kernel: NUL-terminated, pool in CTA order): status, protocol, bit length (MS / MU) and payload bytes of every
    hit against the oracle's arrays; returns '' or a description."""
    import numpy as np

    out, hits = res.out, res.hits
    if not np.array_equal(out["status"], ora_status):
        return f"status differs for {int((out['status'] != ora_status).sum())} messages"
    nh = out["nhits"].astype(np.int64)
    total = int(nh.sum())
    if total != len(hits):
        return f"sum(nhits)={total} != len(hits)={len(hits)}"
    within = np.arange(total) - np.repeat(np.cumsum(nh) - nh, nh)
    order = np.repeat(out["hit_off"].astype(np.int64), nh) + within
    g = hits[order]
    keep = ~(((g["flags"] & 2) != 0) & (g["aux"] != 0))          # continuation rows of a TFA list carry no string
    g = g[keep]
    o = g["off"].astype(np.int64)
    if len(g) != len(ora_hits):
        return f"{len(g)} device hits vs {len(ora_hits)} oracle hits"
    if not np.array_equal(g["proto"].astype(np.int64), ora_hits["proto"].astype(np.int64)):
        return "protocol ids differ"
    if batch.kind <= 1 and not np.array_equal(g["nbits"].astype(np.int64), ora_hits["bit_length"].astype(np.int64)):
        return "bit_length differs"
    pool = np.asarray(pool, dtype=np.uint8)
    nul = np.flatnonzero(pool == 0)
    lens = nul[np.searchsorted(nul, o)] - o
    want = ora_hits["payload_len"].astype(np.int64)
    if not np.array_equal(lens, want):
        return f"payload lengths differ at {int((lens != want).sum())} hits"
    idx = np.repeat(o - (np.cumsum(lens) - lens), lens) + np.arange(int(lens.sum()))
    if pool[idx].tobytes() != ora_pool:
        return "payload bytes differ"
    return ""
