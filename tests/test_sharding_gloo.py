"""N > 1 host logic on CPU: world_size-2 gloo, contiguous shards, max-over-ranks, histogram gather."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _worker(rank, world, port, q):
    sys.path.insert(0, str(ROOT))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from corpus.corpus import Corpus
    from oracle.oracle import Oracle
    from pysignalduino_b200 import pack
    from pysignalduino_b200.protocol_data import load_protocol_table

    protocols = load_protocol_table()
    corp, ora = Corpus(protocols), Oracle(protocols)
    per = 600
    shard = corp.pulse(pack.KIND_MU, world * per, lo=rank * per, hi=(rank + 1) * per)     # bench.py's sharding rule
    status, hits, _pool = ora.run_pulse_raw(shard)
    hist = torch.zeros(len(protocols), dtype=torch.int64)
    for p in hits["proto"]:
        hist[int(p)] += 1
    dist.all_reduce(hist)                      # the only (off-path) collective: per-protocol hit histogram
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)   # timing rule: max over ranks
    if rank == 0:
        whole = corp.pulse(pack.KIND_MU, world * per)
        _, hits_all, _ = ora.run_pulse_raw(whole)
        ref = np.bincount(hits_all["proto"], minlength=len(protocols))
        q.put((hist.numpy().tolist() == ref.tolist(), float(t.item())))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_whole_corpus():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok, tmax = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok and tmax == 2.0


def test_shard_counts_cover_the_corpus():
    sys.path.insert(0, str(ROOT))
    import bench

    for m in (1, 7, 1000, 10_000_000):
        c = bench.shard_counts(m)
        assert sum(c) == m and all(x >= 0 for x in c)
