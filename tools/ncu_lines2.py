#!/usr/bin/env python3
"""Per-source-line warp instructions / samples / entry counts of ONE kernel section of an ncu source page.

    ncu -i X.ncu-rep --page source --csv > src.csv
    cuobjdump -xelf all libsdb200.so; nvdisasm -g -c sdb_pulse.sm_100a.cubin > pulse.sass
    python tools/ncu_lines2.py src.csv pulse.sass <mangled-kernel-substring> <section-index> [lo-hi]
"""
import csv, re, sys
from collections import defaultdict
src_csv, sass, kern, si = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
lo, hi = (map(int, sys.argv[5].split("-")) if len(sys.argv) > 5 else (0, 10**9))
rows = list(csv.reader(open(src_csv)))
secs = []
for r in rows:
    if r and r[0] == "Kernel Name":
        secs.append({"name": r[1], "rows": []}); continue
    if secs: secs[-1]["rows"].append(r)
sec = secs[si]
hdr = sec["rows"][0]; ci = {n: i for i, n in enumerate(hdr)}
line_of = {}; cur = None; active = False
for ln in open(sass, errors="replace"):
    m = re.match(r"\s*//## File \"([^\"]+)\", line (\d+)", ln)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if ln.startswith(".text."): active = kern in ln; continue
    m = re.match(r"\s*/\*([0-9a-f]{4,6})\*/", ln)
    if m and active: line_of[int(m.group(1), 16)] = cur
inst = defaultdict(int); samp = defaultdict(int); mx = defaultdict(int); base = None
for r in sec["rows"][1:]:
    if len(r) < len(hdr): continue
    try: a = int(r[ci["Address"]], 16)
    except ValueError: continue
    if base is None: base = a
    k = line_of.get(a - base, ("?", 0))
    n = int(r[ci["Instructions Executed"]] or 0)
    inst[k] += n; samp[k] += int(r[ci["# Samples"]] or 0); mx[k] = max(mx[k], n)
tot = sum(inst.values())
print(sec["name"], "total warp inst", tot)
for k in sorted(inst):
    if lo <= k[1] <= hi:
        print(f"{k[0]}:{k[1]:5d} inst {inst[k]:11d} {100*inst[k]/tot:5.1f}%  maxcount {mx[k]:9d} samples {samp[k]}")
