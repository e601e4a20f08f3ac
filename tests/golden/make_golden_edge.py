#!/usr/bin/env python3
"""Adversarial fixtures that reach the ACCEPT paths (build container only): edge_ms.json.gz, edge_mu.json.gz.

The round-1 fuzz set only ever produced MS rejects (0 hits in 1 500 records), so tolerance edges, candidate ordering and
the duplicate-id rule were only pinned on the friendly corpus.  Here every record starts from a corpus frame the REAL
reference decodes and is then pushed to the edges of what it accepts:
  * every pulse value is rescaled so that its normalised value lands on / just inside / just outside the accept interval of
    the template value it plays (pattern_utils.py:15-26, :74-76 — incl. the 17 intervals that differ from decimal
    arithmetic);
  * near-duplicate pattern slots are added (several candidates per value: gap ordering :83, cartesian product :111, the
    one-id-for-two-values rule :114), with ids before and after the original in dict order;
  * pattern slots are permuted; the clock pulse (MS) is moved to the 30 % gate (message_synced.py:83-88);
  * D is cut inside / right after the sync, digits are flipped, the last symbol is truncated (reconstructBit).
"""
import gzip
import json
import random
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))
from corpus.corpus import Corpus, batch_to_dicts  # noqa: E402
from oracle import ref_import  # noqa: E402
from pysignalduino_b200 import pack  # noqa: E402
from pysignalduino_b200.protocol_data import load_protocol_table  # noqa: E402
from pysignalduino_b200.table import calculate_tolerance  # noqa: E402


def perturb(m0, typ, rng, protocols):
    m = dict(m0)
    pk = [k for k in m if k[0] == "P" and k[1:].isdigit()]
    if not pk:
        return m
    ops = rng.sample(range(7), rng.randint(1, 3))
    clock = abs(float(m[f"P{int(m['CP'])}"])) if typ == "MS" and f"P{int(m.get('CP', '0'))}" in m else None
    if 0 in ops:                                   # push one value to a tolerance edge of the template value it is closest to
        k = rng.choice(pk)
        v = float(m[k])
        c = clock or rng.choice([float(p["clockabs"]) for p in protocols.values() if "clockabs" in p and float(p["clockabs"]) > 0])
        s = round(v / c) or (1 if v > 0 else -1)
        tol = calculate_tolerance(s)
        edge = s + rng.choice([-1, 1]) * tol + rng.choice([-0.1, -0.05, 0.0, 0.05, 0.1])
        m[k] = str(int(round(edge * c)))
    if 1 in ops and len(pk) < 8:                   # a near-duplicate slot -> several candidates for one value
        src = rng.choice(pk)
        free = [d for d in "01234567" if f"P{d}" not in m]
        if free:
            nk = f"P{rng.choice(free)}"
            nv = str(int(round(float(m[src]) * rng.choice([1.0, 0.97, 1.03, 0.9, 1.1, 1.2]))) + rng.choice([0, 1, -1]))
            items = list(m.items())
            pos = rng.randrange(0, len(items) + 1)
            items.insert(pos, (nk, nv))
            m = dict(items)
    if 2 in ops:                                   # permute the pattern slots (dict order = candidate tie-break order)
        items = list(m.items())
        pats = [(k, v) for k, v in items if k[0] == "P" and k[1:].isdigit()]
        rng.shuffle(pats)
        it = iter(pats)
        m = dict((next(it) if (k[0] == "P" and k[1:].isdigit()) else (k, v)) for k, v in items)
    if 3 in ops and typ == "MS" and clock:         # the clock pulse at the 30 % gate of some protocol
        pc = rng.choice([float(p["clockabs"]) for p in protocols.values() if "sync" in p and float(p.get("clockabs", 0)) > 0])
        c = pc / rng.choice([0.7, 1.3, 0.7001, 1.2999, 0.69, 1.31])
        m[f"P{int(m['CP'])}"] = str(int(round(c)) * (1 if float(m[f"P{int(m['CP'])}"]) > 0 else -1))
    d = m["data"]
    if 4 in ops and len(d) > 8:                    # cut or corrupt D
        how = rng.randrange(4)
        if how == 0:
            d = d[: rng.randrange(2, len(d))]
        elif how == 1:
            i = rng.randrange(len(d))
            d = d[:i] + rng.choice("0123456789") + d[i + 1:]
        elif how == 2:
            d = d[:-1]
        else:
            i = rng.randrange(len(d))
            d = d[:i] + d[i + 1:]
        m["data"] = d
    if 5 in ops:                                   # scale everything (clock drift): normalised values stay, MU clocks do not
        f = rng.choice([0.85, 0.9, 1.1, 1.15, 1.25, 0.75])
        for k in [k for k in m if k[0] == "P" and k[1:].isdigit()]:
            m[k] = str(int(round(float(m[k]) * f)))
    if 6 in ops and typ == "MS":
        m["SP"] = str(rng.randrange(8))
        if rng.random() < 0.3:
            m["R"] = rng.choice(["0", "255", "12", "1q", ""])
    return m


def main():
    protocols = load_protocol_table()
    ref = ref_import.reference_class()()
    corp = Corpus(protocols)
    rng = random.Random(0xED6E)
    for typ, kind, name in (("MS", pack.KIND_MS, "edge_ms.json.gz"), ("MU", pack.KIND_MU, "edge_mu.json.gz")):
        base = [m for m in batch_to_dicts(corp.pulse(kind, 3000)) if m.get("data")]
        good = [m for m in base if ref_import.ref_demodulate(ref, m, typ)[1]]
        recs = []
        for i in range(2000 if typ == "MS" else 1200):
            m = perturb(rng.choice(good), typ, rng, protocols)
            st, res = ref_import.ref_demodulate(ref, m, typ)
            recs.append({"type": typ, "msg": m, "status": st, "results": [list(r) for r in res]})
        path = HERE / name
        with gzip.GzipFile(path, "wb", mtime=0) as gz:
            gz.write(json.dumps(recs, separators=(",", ":")).encode("utf-8"))
        hit = sum(1 for r in recs if r["results"])
        print(f"{name}: {len(recs)} records, {hit} with hits ({100 * hit / len(recs):.0f} %), {sum(len(r['results']) for r in recs)} hits, "
              f"{sum(1 for r in recs if r['status'] != 'ok')} raised, {path.stat().st_size} bytes")


if __name__ == "__main__":
    main()
