#!/usr/bin/env python3
"""Tiny all-kinds run for compute-sanitizer (memcheck / racecheck): every kernel, small batches."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from corpus.corpus import Corpus
from pysignalduino_b200 import SDProtocols, pack

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
sdp = SDProtocols(device=0, mc_repaired=True)
corp = Corpus(sdp.get_protocol_list())
for kind in (pack.KIND_MS, pack.KIND_MU):
    b = corp.pulse(kind, n)
    r = sdp.demodulate_packed(b)
    print(kind, int(r.counters["hits"]), int(r.counters["raised"]))
for kind in (pack.KIND_MC, pack.KIND_MN):
    b = corp.hexmsgs(kind, n)
    r = sdp.engine().demod_host(b, mc_repaired=True)
    print(kind, int(r.counters["hits"]), int(r.counters["raised"]))
print(sdp.postDemo_lengtnPrefix("x", [1, 0, 1] * 20)[0])
