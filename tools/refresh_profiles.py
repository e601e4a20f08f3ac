#!/usr/bin/env python3
"""Turn the artefacts of a final measurement run (gpurun_out/) into the tracked files under profiles/.

    python tools/refresh_profiles.py <ncu-rep of tools/prof_mu.sh> <messages> <launch list csv> <bench json> <reference json> [tag]

tag (default r2_final) names the files: profiles/<tag>_mu_pass_ncu_summary.txt, <tag>_launches.csv, <tag>_bench.json,
<tag>_bench_reference.json; profiles/traffic.json (read by bench.py for roofline.traffic / roofline.issue.executed) is rewritten.
"""
import collections
import csv
import io
import json
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
rep, N, launches, bench, ref = sys.argv[1], int(sys.argv[2]), sys.argv[3], sys.argv[4], sys.argv[5]
TAG = sys.argv[6] if len(sys.argv) > 6 else "r2_final"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = [
    "Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum.per_cycle_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__sass_average_branch_targets_threads_uniform.pct",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__inst_executed_pipe_alu.sum", "smsp__inst_executed_pipe_lsu.sum", "smsp__inst_executed_pipe_fp64.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]
out = ["ncu --set full --clock-control none --import-source on -k 'regex:resolve_kernel|mu_match|mu_emit|scan_kernel' -s 6 -c 6, "
       f"tools/profile_run.py MU {N} 3",
       f"({TAG} build; {N} MU corpus messages, second MU pass: resolve_kernel<MU> -> mu_match_kernel -> mu_emit_kernel -> "
       "scan_kernel<MU> = fused fallback -> sdb_long::resolve_kernel<MU> -> sdb_long::scan_kernel<MU> = the two long-message kernels,",
       " which find an empty list for this corpus)", ""]
tot_inst = tot_dram = tot_t = 0.0
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
for r in rows[2:]:
    d = {h: (v, u) for h, u, v in zip(hdr, units, r)}
    for k in keys:
        if k in d:
            out.append(f"{k:86s}{d[k][0]} {d[k][1]}")
    out.append("")
    tot_inst += float(d["smsp__inst_executed.sum"][0].replace(",", ""))
    for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        tot_dram += float(d[k][0].replace(",", "")) * scale[d[k][1]]
    v, u = d["gpu__time_duration.sum"]
    tot_t += float(v.replace(",", "")) * {"us": 1e-6, "ms": 1e-3, "ns": 1e-9}[u]
out.append(f"sum over the MU pass: {tot_inst / N:.0f} warp instructions / message, {tot_dram / N:.1f} DRAM bytes / message, "
           f"{tot_t * 1e3:.3f} ms for {N} messages")
(ROOT / "profiles" / f"{TAG}_mu_pass_ncu_summary.txt").write_text("\n".join(out) + "\n")
print(out[-1])
src = (f"profiles/{TAG}_mu_pass_ncu_summary.txt (resolve_kernel<MU> + mu_match_kernel + mu_emit_kernel + fallback + long kernels, "
       "dram__bytes_read.sum + dram__bytes_write.sum, smsp__inst_executed.sum)")
tj = {"messages": N, "dram_bytes_per_message": tot_dram / N, "warp_instructions_per_message": tot_inst / N, "source": src}
(ROOT / "profiles" / "traffic.json").write_text(json.dumps({"MU": tj}, indent=1))

rows = list(csv.reader(ln for ln in open(launches) if not ln.startswith("==")))
h = rows[0]
ik, iv, iu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    if len(r) <= iv:
        continue
    v = float(r[iv].replace(",", "")) * {"ms": 1000, "us": 1, "ns": 0.001}.get(r[iu], 1)
    a = agg.setdefault(r[ik], [0, 0.0])
    a[0] += 1
    a[1] += v
T = sum(a[1] for a in agg.values())
for k, (n, t) in agg.items():
    print(f"{k[:56]:56s} launches {n:4d} total {t / 1000:9.3f} ms share {100 * t / T:5.1f}%")
shutil.copy(launches, ROOT / "profiles" / f"{TAG}_launches.csv")
shutil.copy(ref, ROOT / "profiles" / f"{TAG}_bench_reference.json")

d = json.load(open(bench))          # the bench line was produced with the previous traffic.json: same formulas, new counts
r = d["roofline"]
n, ms = d["per_kernel"]["MU"]["messages"], r["avg_launch_ms"]
r["traffic"] = tj["dram_bytes_per_message"] * n
if r.get("issue"):
    wi = tj["warp_instructions_per_message"] * n / (ms / 1e3)
    r["issue"]["executed"] = {"warp_instructions_per_message": tj["warp_instructions_per_message"],
                              "issue_slot_utilisation": wi / (148 * 4 * d["clocks"]["sm_mhz"] * 1e6), "source": src}
(ROOT / "profiles" / f"{TAG}_bench.json").write_text(json.dumps(d))
print("issue (algorithmic) frac", round(r["issue"]["frac"], 3) if r.get("issue") else None, "hbm GB/s", round(r["achieved"], 1), "frac", r["frac"],
      "MU ms", round(ms, 2))
