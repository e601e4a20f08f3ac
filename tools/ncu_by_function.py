#!/usr/bin/env python3
"""Warp instructions per message and SOURCE function of one kernel: the per-line table of tools/ncu_lines2.py summed over the
function each source line belongs to (inlined code counts for the function it was written in, whatever it was inlined into).

    python tools/ncu_lines2.py src.csv pulse.sass <kernel> <section> > lines.txt
    python tools/ncu_by_function.py lines.txt <messages> [source file]
"""
import re
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
lines_file, n = sys.argv[1], float(sys.argv[2])
src_path = Path(sys.argv[3]) if len(sys.argv) > 3 else ROOT / "pysignalduino_b200" / "csrc" / "sdb_pulse.cu"
DEF = re.compile(r"^(?:template <[^>]*>\s*)?(?:__device__|__global__|static|CHAIN_FN|FN_ONE_SITE|int |size_t |unsigned |void )[^;]*?\b(\w+)\(")
marks = []
for i, line in enumerate(src_path.read_text().split("\n"), 1):
    m = DEF.match(line)
    if m and not line.strip().endswith(";") and m.group(1) not in ("__launch_bounds__",):
        marks.append((i, m.group(1)))
    elif line.startswith("__global__") or "__launch_bounds__" in line and line.startswith("__global__"):
        mm = re.search(r"\)\s*(\w+)\(", line)
        if mm:
            marks.append((i, mm.group(1) + " (kernel body)"))
for i, line in enumerate(src_path.read_text().split("\n"), 1):
    mm = re.match(r"^__global__ void __launch_bounds__\([^)]*\)\s*(\w+)\(", line)
    if mm:
        marks = [(a, b) for a, b in marks if a != i] + [(i, mm.group(1) + " (kernel body)")]
marks.sort()


def fn(line_no: int) -> str:
    name = "?"
    for s, nm in marks:
        if s <= line_no:
            name = nm
        else:
            break
    return name


agg, other, tot = {}, {}, 0
for ln in open(lines_file):
    m = re.match(r"(\S+):\s*(\d+) inst\s+(\d+)", ln)
    if not m:
        continue
    f, line_no, c = m.group(1), int(m.group(2)), int(m.group(3))
    tot += c
    if f == src_path.name:
        agg[fn(line_no)] = agg.get(fn(line_no), 0) + c
    else:
        other[f] = other.get(f, 0) + c
for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:18]:
    print(f"{k:34s} {v / n:8.0f} inst/msg {100 * v / tot:5.1f}%")
for k, v in sorted(other.items(), key=lambda kv: -kv[1])[:4]:
    print(f"{'(' + k + ')':34s} {v / n:8.0f} inst/msg {100 * v / tot:5.1f}%")
print(f"total {tot / n:.0f} warp instructions per message")
