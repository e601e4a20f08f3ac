#!/usr/bin/env python3
"""Golden vectors for the scalar pieces (build container only): postDemo_* and helper methods.

Inputs: every 0/1 list literal of the reference's tests/test_postdemodulation.py (extracted with ast — data,
not code), plus seeded mutations (bit flips, truncations, extra leading zeros, random lists).  Every list is
run through ALL nine reference postDemo_* methods, so accept and reject paths are both pinned.
"""
import ast
import gzip
import json
import random
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))
from oracle import ref_import  # noqa: E402

METHODS = ["postDemo_EM", "postDemo_Revolt", "postDemo_FS20", "postDemo_FHT80", "postDemo_FHT80TF",
           "postDemo_WS2000", "postDemo_WS7035", "postDemo_WS7053", "postDemo_lengtnPrefix"]


def bit_lists(path):
    tree = ast.parse(Path(path).read_text(encoding="utf-8"))
    out = []
    for node in ast.walk(tree):
        if isinstance(node, ast.List) and len(node.elts) >= 24 and all(isinstance(e, ast.Constant) and e.value in (0, 1) for e in node.elts):
            out.append([int(e.value) for e in node.elts])
        if isinstance(node, ast.BinOp) and isinstance(node.op, ast.Mult) and isinstance(node.left, ast.List) and isinstance(node.right, ast.Constant):
            if len(node.left.elts) == 1 and isinstance(node.left.elts[0], ast.Constant) and node.left.elts[0].value in (0, 1):
                out.append([int(node.left.elts[0].value)] * int(node.right.value))
    return out


def main():
    ref = ref_import.reference_class()()
    base = bit_lists(ref_import.REFERENCE_ROOT / "tests" / "test_postdemodulation.py")
    rng = random.Random(9)
    cases = [b for b in base]
    for b in base:
        for _ in range(12):
            m = list(b)
            op = rng.randrange(5)
            if op == 0 and m:
                m[rng.randrange(len(m))] ^= 1
            elif op == 1 and len(m) > 2:
                m = m[: rng.randrange(1, len(m))]
            elif op == 2:
                m = [0] * rng.randrange(1, 12) + m
            elif op == 3:
                m = m + [rng.randrange(2) for _ in range(rng.randrange(1, 6))]
            else:
                m = m[rng.randrange(0, min(12, len(m))):]
            cases.append(m)
    for n in (0, 1, 5, 31, 32, 33, 44, 45, 46, 54, 55, 89, 96, 99, 120, 260):
        cases.append([rng.randrange(2) for _ in range(n)])
        cases.append([0] * n)
        cases.append([0] * max(0, n - 1) + [1] * min(1, n))
    recs = []
    for bits in cases:
        for m in METHODS:
            try:
                rc, out = getattr(ref, m)("golden", list(bits))
                recs.append({"method": m, "bits": bits, "rc": int(rc), "out": out})
            except Exception as e:  # noqa: BLE001
                recs.append({"method": m, "bits": bits, "rc": type(e).__name__, "out": None})
    with gzip.GzipFile(HERE / "postdemod.json.gz", "wb", mtime=0) as gz:
        gz.write(json.dumps(recs, separators=(",", ":")).encode())
    acc = sum(1 for r in recs if r["rc"] == 1)
    exc = sum(1 for r in recs if isinstance(r["rc"], str))
    print(f"postdemod.json.gz: {len(recs)} calls, {acc} accepted, {exc} raised")


if __name__ == "__main__":
    main()
