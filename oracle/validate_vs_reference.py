#!/usr/bin/env python3
"""Pin the C oracle against the REAL reference (build container only).

Runs `SDProtocols.demodulate` of /root/reference and the C oracle on the same inputs —
synthetic corpora of the BASELINE shapes plus adversarial fuzz — and compares the canonical
per-message results `(status, [(protocol_id, payload, bit_length)...])`.

    python oracle/validate_vs_reference.py --n 20000 --procs 8
"""
from __future__ import annotations

import argparse
import multiprocessing as mp
import random
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

from corpus.corpus import Corpus, batch_to_dicts  # noqa: E402
from oracle import ref_import  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402
from pysignalduino_b200 import pack  # noqa: E402
from pysignalduino_b200.protocol_data import load_protocol_table  # noqa: E402

_ref = None


def _init():
    global _ref
    _ref = ref_import.reference_class()()


def _work(args):
    msgs, typ = args
    return [ref_import.ref_demodulate(_ref, m, typ) for m in msgs]


def reference_results(msgs, typ, procs):
    chunk = max(1, len(msgs) // (procs * 8))
    parts = [(msgs[i : i + chunk], typ) for i in range(0, len(msgs), chunk)]
    with mp.Pool(procs, initializer=_init) as pool:
        out = []
        for r in pool.imap(_work, parts):
            out.extend(r)
    return out


def fuzz_pulse(rng: random.Random, n: int, kind: int, protocols):
    """Adversarial messages: few pattern ids, random digit soups, near-tolerance pulse values."""
    out = []
    ids = list(protocols)
    mu_ids = [k for k, v in protocols.items() if "clockabs" in v and "one" in v]
    for _ in range(n):
        pr = protocols[rng.choice(mu_ids)]
        clock = float(pr["clockabs"])
        if clock <= 0:
            clock = rng.randint(250, 600)
        vals = []
        for key in ("start", "sync", "one", "zero", "float"):
            v = pr.get(key)
            if isinstance(v, list):
                vals.extend(float(x) for x in v)
        vals = list(dict.fromkeys(vals))[: rng.randint(2, 8)]
        npat = min(8, max(2, len(vals) + rng.randint(0, 2)))
        slots = rng.sample(range(8), npat)
        m = {}
        for s in slots:
            if vals and rng.random() < 0.8:
                v = rng.choice(vals)
                # edge-of-tolerance jitter
                j = rng.choice([1.0, 1.0, 0.7, 0.82, 1.18, 1.3, 0.95, 1.05, rng.uniform(0.6, 1.4)])
                m[f"P{s}"] = str(int(round(v * clock * j)) + rng.choice([0, 0, 0, 1, -1, 5, -5]))
            else:
                m[f"P{s}"] = str(rng.randint(-9000, 9000))
        L = rng.randint(1, 400)
        mode = rng.random()
        if mode < 0.4:
            d = "".join(str(rng.choice(slots)) for _ in range(L))
        elif mode < 0.8:
            a, b, c = (str(rng.choice(slots)) for _ in range(3))
            syms = [a + b, a + c, b + a, c + a]
            d = "".join(rng.choice(syms[: rng.randint(1, 4)]) for _ in range(L // 2))
            if rng.random() < 0.5:
                pre = "".join(str(rng.choice(slots)) for _ in range(rng.randint(0, 8)))
                d = pre + d
            if rng.random() < 0.3:
                k = rng.randint(0, len(d))
                d = d[:k] + str(rng.choice(slots)) + d[k:]
        else:
            a, b = str(rng.choice(slots)), str(rng.choice(slots))
            d = (a + b) * (L // 2) + rng.choice(["", a, b])
        if not d:
            d = str(slots[0])
        m["data"] = d[:1024]
        if kind == pack.KIND_MS:
            near = [s_ for s_ in slots if 0.6 * clock <= float(m[f"P{s_}"]) <= 1.4 * clock]
            m["CP"] = str(rng.choice(near) if near and rng.random() < 0.9 else rng.choice(slots))
            m["SP"] = str(rng.choice(slots))
        if rng.random() < 0.5:
            m["R"] = str(rng.randint(0, 255))
        out.append(m)
    return out


def compare(name, msgs, kind, typ, ora, procs):
    t0 = time.time()
    exp = reference_results(msgs, typ, procs)
    t1 = time.time()
    got = ora.run_pulse(pack.pack_pulse(msgs, kind), nthreads=procs)
    t2 = time.time()
    bad = [i for i, (e, g) in enumerate(zip(exp, got)) if e != g]
    nh = sum(len(e[1]) for e in exp)
    nr = sum(1 for e in exp if e[0] != "ok")
    print(f"{name}: n={len(msgs)} hits={nh} raised={nr} mismatches={len(bad)} "
          f"(reference {t1 - t0:.1f}s = {len(msgs) / (t1 - t0):.0f} msg/s on {procs} procs, oracle {t2 - t1:.2f}s)")
    for i in bad[:5]:
        print("  MISMATCH", i, msgs[i])
        print("    ref   :", exp[i])
        print("    oracle:", got[i])
    return len(bad)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=20000)
    ap.add_argument("--procs", type=int, default=8)
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args()
    protocols = load_protocol_table()
    for pid, pr in protocols.items():
        pr.setdefault("active", True)
        pr.setdefault("name", f"Protocol_{pid}")
    ora = Oracle(protocols)
    corp = Corpus(protocols)
    rng = random.Random(a.seed)
    bad = 0
    bad += compare("corpus MS", batch_to_dicts(corp.pulse(pack.KIND_MS, a.n)), pack.KIND_MS, "MS", ora, a.procs)
    bad += compare("corpus MU", batch_to_dicts(corp.pulse(pack.KIND_MU, a.n)), pack.KIND_MU, "MU", ora, a.procs)
    bad += compare("fuzz MS", fuzz_pulse(rng, a.n, pack.KIND_MS, protocols), pack.KIND_MS, "MS", ora, a.procs)
    bad += compare("fuzz MU", fuzz_pulse(rng, a.n, pack.KIND_MU, protocols), pack.KIND_MU, "MU", ora, a.procs)
    print("TOTAL MISMATCHES", bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
