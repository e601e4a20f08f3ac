#!/usr/bin/env python3
"""Golden vectors for direct mcBit2* / mcRaw / mcraw calls (build container only).

Inputs: every 0/1 string literal of the reference's tests/test_manchester_protocols.py (extracted with ast —
data, not code; adjacent literals are concatenated by the parser), seeded mutations of them, and random strings
with the decoders' sync words planted.  Every string is run through ALL thirteen reference decoders with
 * the table ids that name the decoder, an id that is not in the table, an int id (119 / 58),
 * the protocol patches the reference's own tests apply (`proto._protocols[pid] = {...}`),
 * mcbitnum = len(bits), None, and off-by-some values,
so accept, every reject message and the raising paths are pinned.
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

METHODS = ["mcBit2Funkbus", "mcBit2Sainlogic", "mcBit2AS", "mcBit2Hideki", "mcBit2Maverick", "mcBit2OSV1",
           "mcBit2OSV2o3", "mcBit2OSPIR", "mcRaw", "mcraw", "mcBit2TFA", "mcBit2Grothe", "mcBit2SomfyRTS"]

# the patches of tests/test_manchester_protocols.py:211-400 (values only)
PATCHES = [
    None,
    {"5058": {"length_min": 51, "length_max": 52, "name": "Unittest TFA"}},
    {"5058": {"length_min": 52, "length_max": 52, "name": "Unittest TFA"}},
    {"5058": {"length_min": 80, "length_max": 100, "name": "Unittest TFA"}},
    {"5058": {"length_min": 51, "length_max": 100, "name": "Unittest TFA"}},
    {"5043": {"length_min": 52, "length_max": 100, "name": "Unittest AS"}},
    {"5043": {"length_min": 52, "length_max": 60, "name": "Unittest AS"}},
    {"119": {"length_min": 50, "name": "TestLength"}},
]


def bit_strings(path):
    tree = ast.parse(Path(path).read_text(encoding="utf-8"))
    out = []
    for node in ast.walk(tree):
        if isinstance(node, ast.Constant) and isinstance(node.value, str) and len(node.value) >= 16 and set(node.value) <= {"0", "1"}:
            out.append(node.value)
    return sorted(set(out), key=lambda s: (len(s), s))


def call(ref, method, bits, pid, mcbitnum):
    try:
        if method == "mcraw":
            rc, out = ref.mcraw("golden", bits, pid, mcbitnum)
        else:
            rc, out = getattr(ref, method)("golden", bits, pid, mcbitnum)
        return int(rc), out
    except Exception as e:  # noqa: BLE001
        return type(e).__name__, None


def main():
    Ref = ref_import.reference_class()
    base = bit_strings(ref_import.REFERENCE_ROOT / "tests" / "test_manchester_protocols.py")
    rng = random.Random(0x3C)
    cases = list(base)
    for b in base:
        for _ in range(6):
            m = list(b)
            op = rng.randrange(5)
            if op == 0:
                m[rng.randrange(len(m))] = "10"[int(m[rng.randrange(len(m))])]
            elif op == 1:
                m = m[: rng.randrange(1, len(m))]
            elif op == 2:
                m = ["1"] * rng.randrange(1, 12) + m
            elif op == 3:
                m = m + [rng.choice("01") for _ in range(rng.randrange(1, 9))]
            else:
                m = m[rng.randrange(0, min(12, len(m))):]
            cases.append("".join(m))
    syncs = ["01100", "010100", "1100", "111111111101", "1111111111101", "1101"]
    for n in (0, 1, 4, 15, 16, 20, 31, 32, 33, 48, 55, 56, 57, 58, 64, 100, 127, 128, 129, 200, 260):
        cases.append("".join(rng.choice("01") for _ in range(n)))
        cases.append("0" * n)
        cases.append("1" * n)
        if n >= 20:
            s = ["01"[rng.random() < 0.3] for _ in range(n)]
            for _ in range(3):
                w = rng.choice(syncs)
                p = rng.randrange(0, n - len(w))
                s[p : p + len(w)] = list(w)
            cases.append("".join(s))
    # TFA: a frame repeated 2-3 times behind its sync words, optionally one corrupted copy
    for _ in range(24):
        ln = rng.randrange(30, 70)
        frame = "".join(rng.choice("01") for _ in range(ln)).replace("11111", "11011")
        reps = rng.randrange(2, 4)
        s = "1111111111101" + frame
        for r in range(1, reps):
            f2 = frame if rng.random() < 0.8 else frame[:-1] + "10"[int(frame[-1])]
            s += "1111111111101" + f2
        cases.append(s + "".join(rng.choice("01") for _ in range(rng.randrange(0, 6))))
    cases = sorted(set(cases), key=lambda s: (len(s), s))

    groups = []
    total = acc = exc = 0
    for patch in PATCHES:
        ref = Ref()
        if patch:
            for pid, row in patch.items():
                ref._protocols[pid] = dict(row)
        table = ref._protocols
        recs = []
        for bits in cases:
            for m in METHODS:
                tail = m if m != "mcraw" else "mcraw"
                if patch:
                    ids = list(patch.keys())
                else:
                    ids = [pid for pid, pr in table.items() if str(pr.get("method", "")).split(".")[-1] == tail][:3]
                    ids += ["9999"]
                    if m == "mcBit2Funkbus":
                        ids += [119]
                    if m == "mcBit2Hideki":
                        ids += ["12", 58]
                    if not ids[:-1]:
                        ids += ["10"]
                for pid in ids:
                    nums = [len(bits), None]
                    if not patch:
                        nums += [len(bits) + 1, max(0, len(bits) - 3), 32, 57, 128]
                    for mcbitnum in dict.fromkeys(nums):
                        rc, out = call(ref, m, bits, pid, mcbitnum)
                        recs.append([m, bits, pid, mcbitnum, rc, out])
                        total += 1
                        acc += rc == 1
                        exc += isinstance(rc, str)
        groups.append({"patch": patch, "calls": recs})
    with gzip.GzipFile(HERE / "mc_units.json.gz", "wb", mtime=0) as gz:
        gz.write(json.dumps(groups, separators=(",", ":")).encode())
    print(f"mc_units.json.gz: {len(cases)} strings, {total} calls, {acc} accepted, {exc} raised")


if __name__ == "__main__":
    main()
