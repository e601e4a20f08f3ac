#!/usr/bin/env python3
"""Golden vectors for direct Conv*(msg_data) calls (helpers.py:223-716), build container only: conv_units.json.gz.

Inputs: every upper-case hex literal (>= 10 characters) of the reference's tests/test_helpers.py and
tests/test_mn_bresser_lightning.py (ast: data, not code), seeded mutations of them, and the first 500 messages of
the MN corpus.  Each one goes through ALL seven reference converters with an int protocol id, as the reference's
tests call them.
"""
import ast
import gzip
import json
import random
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))
from corpus.corpus import Corpus  # noqa: E402
from oracle import ref_import  # noqa: E402
from pysignalduino_b200 import pack  # noqa: E402
from pysignalduino_b200.protocol_data import load_protocol_table  # noqa: E402

METHODS = ["ConvBresser_lightning", "ConvBresser_5in1", "ConvBresser_6in1", "ConvBresser_7in1", "ConvPCA301",
           "ConvKoppFreeControl", "ConvLaCrosse"]
HEX = set("0123456789ABCDEF")


def hex_literals(path):
    out = []
    for node in ast.walk(ast.parse(Path(path).read_text(encoding="utf-8"))):
        if isinstance(node, ast.Constant) and isinstance(node.value, str) and 10 <= len(node.value) <= 512 and set(node.value) <= HEX:
            out.append(node.value)
    return out


def main():
    ref = ref_import.reference_class()()
    rng = random.Random(0xC0)
    base = []
    for f in ("test_helpers.py", "test_mn_bresser_lightning.py", "test_mn_parser.py"):
        base += hex_literals(ref_import.REFERENCE_ROOT / "tests" / f)
    base = sorted(set(base))
    cases = list(base)
    for b in base:
        for _ in range(4):
            m = list(b)
            op = rng.randrange(3)
            if op == 0:
                m[rng.randrange(len(m))] = rng.choice("0123456789ABCDEF")
            elif op == 1:
                m = m[: rng.randrange(1, len(m))]
            else:
                m = m + [rng.choice("0123456789ABCDEF") for _ in range(rng.randrange(1, 5))]
            cases.append("".join(m))
    corp = Corpus(load_protocol_table())
    b = corp.hexmsgs(pack.KIND_MN, 500)
    for i in range(b.n):
        d = pack.unpack_hex(b, i).get("data")
        if d:
            cases.append(d)
    cases += ["", "0", "AA"]
    recs = []
    nacc = 0
    for data in cases:
        for k, m in enumerate(METHODS):
            msg = {"data": data, "protocol_id": 100 + k}
            try:
                out = getattr(ref, m)(dict(msg))
                recs.append([m, msg, "ok", out])
                nacc += bool(out)
            except Exception as e:  # noqa: BLE001
                recs.append([m, msg, type(e).__name__, None])
    with gzip.GzipFile(HERE / "conv_units.json.gz", "wb", mtime=0) as gz:
        gz.write(json.dumps(recs, separators=(",", ":")).encode())
    print(f"conv_units.json.gz: {len(cases)} strings, {len(recs)} calls, {nacc} accepted, "
          f"{sum(1 for r in recs if r[2] != 'ok')} raised")


if __name__ == "__main__":
    main()
