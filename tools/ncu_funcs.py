#!/usr/bin/env python3
"""Samples / warp instructions per device function (noinline sub-functions) of one kernel.
    python tools/ncu_funcs.py src.csv pulse.sass <kernel-substring>"""
import csv, re, sys
from collections import defaultdict
src_csv, sass, kern = sys.argv[1:4]
func_of = {}
cur = None
active = False
for ln in open(sass, errors="replace"):
    if ln.startswith(".text."):
        active = kern in ln
        cur = "<kernel body>"
        continue
    m = re.match(r"^(\$?[\w$]+):\s*$", ln.strip())
    if m and active and not m.group(1).startswith(".L"):
        name = m.group(1)
        mm = re.findall(r"\$_ZN3sdb\d+(\w+?)E", name)
        cur = mm[-1] if mm else name[-40:]
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,6})\*/", ln)
    if m and active:
        func_of[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; ci = {n: i for i, n in enumerate(hdr)}
samp = defaultdict(int); inst = defaultdict(int); base = None; ts = ti = 0
for r in rows[hi + 1:]:
    if len(r) < len(hdr): continue
    try: addr = int(r[ci["Address"]], 16)
    except ValueError: continue
    if base is None: base = addr
    f = func_of.get(addr - base, "?")
    s = int(float(r[ci["# Samples"]] or 0)); n = int(float(r[ci["Instructions Executed"]] or 0))
    samp[f] += s; inst[f] += n; ts += s; ti += n
print(f"total samples {ts}, warp instructions {ti}")
for k, v in sorted(inst.items(), key=lambda kv: -kv[1]):
    print(f"{v:12d} inst {100*v/ti:5.1f}%   samples {samp[k]:8d} {100*samp[k]/ts:5.1f}%   {k}")
