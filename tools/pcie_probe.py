#!/usr/bin/env python3
"""Host <-> device copy bandwidth of every rank at once (no demodulator code involved): isolates what limits the end-to-end
path when N ranks share one host — PCIe / host memory / NUMA placement.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_probe.py [--numa]

Each rank copies a 1 GiB pinned buffer H2D, D2H and both at once, 10 times, after a barrier; rank 0 prints one JSON line
with min / mean GB/s over the ranks.  --numa pins the rank to its GPU's NVML CPU affinity before the buffers are allocated
(what bench.py does)."""
import json
import os
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    cpus = None
    if "--numa" in sys.argv:
        from pysignalduino_b200.capi import bind_to_gpu_numa

        cpus = bind_to_gpu_numa(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = 1 << 30
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_in.fill_(1)
    d_a = torch.empty(n, dtype=torch.uint8, device=dev)
    d_b = torch.ones(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def run(mode, reps=10):
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            if mode in ("h2d", "both"):
                with torch.cuda.stream(s1):
                    d_a.copy_(h_in, non_blocking=True)
            if mode in ("d2h", "both"):
                with torch.cuda.stream(s2):
                    h_out.copy_(d_b, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        return reps * n * (2 if mode == "both" else 1) / dt / 1e9

    out = {}
    for mode in ("h2d", "d2h", "both"):
        run(mode, 2)
        g = torch.tensor([run(mode)], dtype=torch.float64, device=dev)
        if world > 1:
            lst = [torch.zeros_like(g) for _ in range(world)]
            dist.all_gather(lst, g)
            vals = [float(x.item()) for x in lst]
        else:
            vals = [float(g.item())]
        out[mode] = {"min_GBps": min(vals), "mean_GBps": sum(vals) / len(vals), "sum_GBps": sum(vals)}
    if rank == 0:
        print(json.dumps({"ranks": world, "numa_binding": bool(cpus), "cpus_rank0": len(cpus) if cpus else os.cpu_count(), "copy": out}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
