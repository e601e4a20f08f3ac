#!/usr/bin/env python3
"""Golden vectors for MU messages with MANY matches (build container only): crafted_mu.json.gz.

The corpus has 2-4 frames per message; the device path has separate code for survivors with more than 4 matches
(a lane records 4 at a time), for messages with more than 64 matches (fused fallback kernel) and for empty captures
(IndexError).  These messages hit all of them: the same short frame repeated up to the 1024-digit limit, built
from a protocol's own start / one / zero templates, expected results from the reference itself.
"""
import random
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))
sys.path.insert(0, str(HERE))
from make_golden import dump, run  # noqa: E402
from oracle import ref_import  # noqa: E402
from pysignalduino_b200.protocol_data import load_protocol_table  # noqa: E402


def fixed_messages(rng):
    msgs = []
    # protocol-15 family: no start, one [1,-1], zero [1,-2], clock 700, 10..20 symbols
    for nsym, reps in ((10, 46), (11, 42), (12, 30), (20, 24)):
        frames = ["".join(rng.choice(("01", "02")) for _ in range(nsym)) + "03" for _ in range(reps)]
        msgs.append({"P0": "700", "P1": "-700", "P2": "-1400", "P3": "-9000", "data": "".join(frames)[:1024], "R": "12"})
    # [-30,1] start family (ids 30 / 29 / 83 / 81 / 79 / 86): one [-2,1], zero [-1,2]
    for clock, nsym, reps in ((330, 12, 39), (335, 12, 39), (235, 12, 30), (500, 12, 39), (350, 14, 34)):
        frames = ["40" + "".join(rng.choice(("10", "32")) for _ in range(nsym)) for _ in range(reps)]
        msgs.append({"P0": str(clock), "P1": str(-2 * clock), "P2": str(2 * clock), "P3": str(-clock), "P4": str(-30 * clock),
                     "data": "".join(frames)[:1024]})
    # protocol 31 (no length_min, 6-pulse start): many tiny matches, some with an empty capture (IndexError)
    for reps, body in ((60, 3), (80, 1), (70, 0)):
        frames = ["010102" + "".join(rng.choice(("03", "41")) for _ in range(body)) + "5" for _ in range(reps)]
        msgs.append({"P0": "315", "P1": "-284", "P2": "-1197", "P3": "-567", "P4": "630", "P5": "-5000",
                     "data": "".join(frames)[:1024]})
    return msgs


def protocol_messages(rng, protocols, count):
    """Frames of a random MU protocol's own templates, repeated until D is (nearly) full."""
    ids = [pid for pid, pr in protocols.items() if "clockabs" in pr and isinstance(pr.get("one"), list) and isinstance(pr.get("zero"), list)]
    out = []
    while len(out) < count:
        pid = rng.choice(ids)
        pr = protocols[pid]
        clock = float(pr["clockabs"])
        start = pr.get("start") if isinstance(pr.get("start"), list) else []
        vals = []
        for v in list(start) + list(pr["one"]) + list(pr["zero"]):
            if v not in vals:
                vals.append(v)
        if len(vals) > 7:
            continue
        sep = -(abs(max(vals, key=abs)) * 3 + 7)
        slots = vals + [sep]
        rng.shuffle(slots)
        ident = {v: str(i) for i, v in enumerate(slots)}
        lmin = int(pr.get("length_min", 8) or 8)
        nsym = max(1, lmin + rng.choice((-1, 0, 0, 1, 3)))
        frame_len = len(start) + nsym * len(pr["one"]) + 1
        reps = rng.randrange(3, max(4, min(90, 1024 // frame_len + 1)))
        frames = []
        for _ in range(reps):
            f = "".join(ident[v] for v in start)
            for _ in range(nsym):
                f += "".join(ident[v] for v in rng.choice((pr["one"], pr["zero"])))
            frames.append(f + (ident[sep] if rng.random() < 0.8 else ""))
        msg = {f"P{i}": str(int(round(v * clock * rng.uniform(0.97, 1.03)))) for i, v in enumerate(slots)}
        msg["data"] = "".join(frames)[:1024]
        out.append(msg)
    return out


def main():
    protocols = load_protocol_table()
    ref = ref_import.reference_class()()
    rng = random.Random(0xC0FFEE)
    msgs = fixed_messages(rng) + protocol_messages(rng, protocols, 240)
    dump("crafted_mu.json.gz", run(ref, [("MU", m) for m in msgs]))


if __name__ == "__main__":
    main()
