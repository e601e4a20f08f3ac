#!/usr/bin/env python3
"""Golden vectors for framing / decompression only (build container only): frames.json.gz.

6 000 byte-level fuzz cases around the reduced ("Mred=1") payload grammar -> the reference's extract_payload
(signalduino/parser/base.py:174-193).  Lines are stored latin-1 decoded, as the reference's transport delivers them.
"""
import gzip
import json
import random
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))
sys.path.insert(0, str(HERE))
from oracle import ref_import  # noqa: E402
from tests.common import load_golden  # noqa: E402


def main():
    mod = ref_import.reference_parser_module()
    rng = random.Random(0xF4A3)
    seeds = [r["line"] for r in load_golden("lines.json.gz") if r["rfmode"] is None and r["line"].startswith("\x02")]
    reduced = [s for s in seeds if any(ord(c) > 127 for c in s)]
    alphabet = [chr(c) for c in list(range(0x20, 0x7F)) + list(range(0x80, 0x100)) + [0x09, 0x0B, 0x1C, 0x1F, 0x02, 0x03]]
    cases = []
    for _ in range(6000):
        s = list(rng.choice(reduced if rng.random() < 0.8 else seeds))
        for _ in range(rng.randrange(1, 5)):
            op = rng.randrange(6)
            i = rng.randrange(1, max(2, len(s) - 1))
            if op == 0:
                s[i] = rng.choice(alphabet)
            elif op == 1:
                s.insert(i, rng.choice(";;;DdMmoCSRFPp=0123456789ABCDEFabcdef" + "\xb2\xaa\xd7\xf7\xdf\xff\xb5"))
            elif op == 2:
                del s[i]
            elif op == 3:
                s.insert(i, ";" + rng.choice(["C1", "S3", "R2A", "F64", "o5", "m0", "M", "D", "d", "x=1", "\xa5\x81\x82", "Zz", "1A", "\xe9\xe9"]))
            elif op == 4:
                s[i:i] = list(rng.choice(["  ", "\t", "\x1c", "\xa0", "\x85"]))
            else:
                j = rng.randrange(1, max(2, len(s) - 1))
                s[i], s[j] = s[j], s[i]
        line = "".join(s).replace("\n", "")
        if rng.random() < 0.1:
            line = rng.choice([" ", "\t", "\xa0", "\x1d"]) + line + rng.choice(["\r", " ", "\x85"])
        cases.append(line)
    cases = list(dict.fromkeys(cases))
    recs = [[c, mod.base.extract_payload(c)] for c in cases]
    with gzip.GzipFile(HERE / "frames.json.gz", "wb", mtime=0) as gz:
        gz.write(json.dumps(recs, separators=(",", ":")).encode("utf-8"))
    print(f"frames.json.gz: {len(recs)} lines, {sum(1 for r in recs if r[1] is not None)} framed, "
          f"{(HERE / 'frames.json.gz').stat().st_size} bytes")


if __name__ == "__main__":
    main()
