#!/usr/bin/env python3
"""Entry counts and warp instructions per device function of one kernel section of an ncu source page.
    python tools/ncu_entries.py src.csv pulse.sass <mangled-kernel-substring> <section-index>"""
import csv, re, sys
src, sass, kern, si = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
rows = list(csv.reader(open(src)))
secs = []
for r in rows:
    if r and r[0] == "Kernel Name":
        secs.append({"name": r[1], "rows": []}); continue
    if secs: secs[-1]["rows"].append(r)
sec = secs[si]
hdr = sec["rows"][0]; ci = {n: i for i, n in enumerate(hdr)}
func_of = {}; cur = None; active = False
for ln in open(sass, errors="replace"):
    if ".section\t.text." in ln or ln.startswith(".text."):
        active = kern in ln; cur = "<body>"; continue
    m = re.match(r"^(\$?[\w$]+):\s*$", ln.strip())
    if m and active and not m.group(1).startswith(".L"):
        cur = m.group(1); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,6})\*/", ln)
    if m and active: func_of[int(m.group(1), 16)] = cur
base = None; first = {}; tot = {}
for r in sec["rows"][1:]:
    if len(r) < len(hdr): continue
    try: a = int(r[ci["Address"]], 16)
    except ValueError: continue
    if base is None: base = a
    f = func_of.get(a - base, "?")
    n = int(r[ci["Instructions Executed"]] or 0)
    if f not in first: first[f] = n
    tot[f] = tot.get(f, 0) + n
print(sec["name"], "total", sum(tot.values()))
for f in first:
    print(f"   {f[-44:]:44s} entries {first[f]:9d} inst {tot[f]:11d} inst/entry {tot[f] // max(first[f], 1)}")
