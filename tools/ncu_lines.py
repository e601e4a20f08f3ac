#!/usr/bin/env python3
"""Join an ncu SASS source page (csv) with nvdisasm line info: samples / instructions per CUDA source line.

    ncu -i X.ncu-rep --page source --csv > src.csv
    cuobjdump -xelf all libsdb200.so; nvdisasm -g -c sdb_pulse.sm_100a.cubin > pulse.sass
    python tools/ncu_lines.py src.csv pulse.sass <kernel-substring> [top]
"""
import csv
import re
import sys
from collections import defaultdict

src_csv, sass, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40

# address -> (file, line) from nvdisasm
line_of = {}
cur = None
infn = False
for ln in open(sass, errors="replace"):
    if ln.startswith(".text.") or "Function" in ln:
        pass
    m = re.match(r"\s*//## File \"([^\"]+)\", line (\d+)", ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if ln.startswith(".text."):
        infn = kern in ln
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,6})\*/", ln)
    if m and infn:
        line_of[int(m.group(1), 16)] = cur

rows = list(csv.reader(open(src_csv)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ci = {n: i for i, n in enumerate(hdr)}
samp = defaultdict(int)
inst = defaultdict(int)
tot_s = tot_i = 0
base = None
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    try:
        addr = int(r[ci["Address"]], 16)
    except ValueError:
        continue
    if base is None:
        base = addr
    key = line_of.get(addr - base, ("?", 0))
    s = int(float(r[ci["# Samples"]] or 0))
    n = int(float(r[ci["Instructions Executed"]] or 0))
    samp[key] += s
    inst[key] += n
    tot_s += s
    tot_i += n
print(f"total samples {tot_s}, warp instructions {tot_i}")
print("--- by samples")
for k, v in sorted(samp.items(), key=lambda kv: -kv[1])[:top]:
    print(f"{v:9d} {100*v/tot_s:5.1f}%  inst {inst[k]:11d} {100*inst[k]/tot_i:5.1f}%  {k[0]}:{k[1]}")
