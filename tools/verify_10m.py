#!/usr/bin/env python3
"""BASELINE.json target check: bit-exact decode of the 10 M-message mixed corpus (config 5: 40 % MS, 40 % MU, 15 % MC,
5 % MN, x every protocol of each class) against the CPU oracle, every hit compared (status, protocol, bit_length,
payload bytes).  Run on the GPU box:  python tools/verify_10m.py [messages] > gpurun_out/verify_10m.json

The oracle (oracle/, C restatement pinned against the reference) is the checker here, exactly as in tests/.
"""
import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from corpus.corpus import Corpus  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402
from pysignalduino_b200 import SDProtocols, pack  # noqa: E402
from tests.common import compare_payloads, compare_raw  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
MIX = (("MS", 0, 0.40), ("MU", 1, 0.40), ("MC", 2, 0.15), ("MN", 3, 0.05))
STEP = 500_000
threads = os.cpu_count() or 1
sdp = SDProtocols(device=0, mc_repaired=True)
protocols = sdp.get_protocol_list()
corp, ora = Corpus(protocols), Oracle(protocols)
counts = [int(M * f) for _, _, f in MIX]
counts[0] += M - sum(counts)
report = {"messages": M, "threads": threads, "classes": {}, "mismatches": 0}
t_gpu = t_cpu = t_gpu_payload = 0.0
for (name, kind, _), n in zip(MIX, counts):
    hits = raised = 0
    bad = []
    for lo in range(0, n, STEP):
        hi = min(n, lo + STEP)
        b = corp.pulse(kind, n, lo=lo, hi=hi) if kind <= 1 else corp.hexmsgs(kind, n, lo=lo, hi=hi)
        t0 = time.perf_counter()
        res = sdp.engine().demod_host(b, mc_repaired=True)
        t_gpu += time.perf_counter() - t0
        t0 = time.perf_counter()
        if kind <= 1:
            status, ohits, pool = ora.run_pulse_raw(b, nthreads=threads)
        else:
            status, ohits, pool = ora.run_hex_raw(b, mc_repaired=True, nthreads=threads)
        t_cpu += time.perf_counter() - t0
        msg = compare_raw(sdp, b, res, status, ohits, pool, check_bits=kind <= 1)
        if msg:
            bad.append(f"[{lo}:{hi}] {msg}")
        # the same block through sdb_demod_host_payloads: strings written by the device format kernel
        t0 = time.perf_counter()
        res2, dpool = sdp.engine().demod_payloads(b, mc_repaired=True)
        t_gpu_payload += time.perf_counter() - t0
        msg = compare_payloads(b, res2, dpool, status, ohits, pool)
        if msg:
            bad.append(f"[{lo}:{hi}] device-formatted: {msg}")
        hits += len(res.hits)
        raised += int(res.counters["raised"])
    report["classes"][name] = {"messages": n, "hits": hits, "raised": raised, "mismatch": bad}
    report["mismatches"] += len(bad)
    print(f"{name}: {n} messages, {hits} hits, {raised} raised, {'OK' if not bad else bad}", file=sys.stderr)
report["gpu_host_call_s"] = round(t_gpu, 2)
report["gpu_payload_call_s"] = round(t_gpu_payload, 2)
report["checked"] = "status, protocol, bit length and payload bytes of every hit, through sdb_demod_host + sdb_format_hits AND through sdb_demod_host_payloads (device format kernel)"
report["oracle_s"] = round(t_cpu, 2)
report["bit_exact"] = report["mismatches"] == 0
print(json.dumps(report))
sys.exit(0 if report["bit_exact"] else 1)
