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
