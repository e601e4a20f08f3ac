#!/usr/bin/env python3
"""bench.py — messages demodulated / second on the BASELINE.json workload.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--messages M] [--impl ours|reference]

Workload (BASELINE.json configs[4]): a mixed corpus, 40 % MS / 40 % MU /
15 % MC / 5 % MN, M messages per GPU (default 10 M), each message decoded against EVERY protocol of
its class (47 MS / 129 MU protocols; MC / MN name their protocol).  The corpus is partitioned by
message type when it is packed (the type is the first two characters of a firmware line); one
"step" = one pass of the hot path over the whole per-GPU shard (per 1 048 576 resident messages: MS = resolve + scan,
MU = resolve + match + emit + fused fallback; MC and MN one launch each; the host-buffer path pipelines in stages of
262 144 messages).
Rank r of N decodes messages [r*M, (r+1)*M) of the N*M-message corpus: weak scaling, replicated
protocol table, no collective on the decode path.

ours:      `value` = device-resident throughput (inputs already in HBM), CUDA events on the
           launching stream, max over ranks; `e2e` = same metric through the C ABI with HOST buffers
           (pinned H2D + kernels + D2H + the payload string of every hit inside the timed region:
           sdb_demod_host_payloads); `roofline` for the dominant kernel (MU): algorithmic bytes
           48 + ceil(dlen/2) + sum_hits(16 + ceil(nbits/8)) per message against the measured HBM peak, and
           `roofline.issue`: SURVEY §8(d) algorithmic integer ops per message against 148 x 4 x 32 x f_SM.
           `cpu_baseline` (rank 0, N = 1): the REAL Python reference (oracle/_ref, installed by oracle/make_ref.py)
           in Pool(os.cpu_count()) on a stratified sample, compared message by message with the GPU output of the
           same rows (`parity_checked_messages`), next to the C oracle port.
           `--scaling strong` shards ONE --messages corpus over the ranks (BASELINE config 5); the default is weak
           (--messages per GPU) and, for N > 1, a `strong` object measured in the same run.
reference: the reference's own CPU implementation on all host cores: the Python reference from oracle/_ref when it is
           there (kind "reference"), else the C oracle port (kind "port"), on bounded samples of the same corpus.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

MIX = (("MS", 0, 0.40), ("MU", 1, 0.40), ("MC", 2, 0.15), ("MN", 3, 0.05))
METRIC = "messages demodulated/sec x all protocols (bit-exact)"
UNIT = "messages/s"


_REAL_STDOUT = None


def capture_stdout():
    """Library banners (e.g. "NCCL version ...") must not pollute stdout: route fd 1 to stderr and keep the
    real stdout for the ONE JSON line."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line) -> None:
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def shard_counts(m: int):
    c = [int(m * f) for _, _, f in MIX]
    c[0] += m - sum(c)
    return c


def config_dict(args, n_gpus):
    strong = getattr(args, "scaling", "weak") == "strong"
    return {
        "workload": "config5: mixed MS/MU/MC/MN corpus (40/40/15/5 %), seed 0x5D05 family, x all protocols of each class",
        "messages_per_gpu": args.messages // n_gpus if strong else args.messages,
        "total_messages": (args.messages // n_gpus) * n_gpus if strong else args.messages * n_gpus,
        "protocols": {"MS": 47, "MU": 129, "MC": 12, "MN": 8},
        "sharding": "contiguous message ranges per GPU, replicated protocol table, no collective",
        "l2_policy": "inputs larger than L2 (per-step input >> 126 MB); no explicit flush",
        "mc_mode": "repaired",
    }


# --------------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting"}

    def __init__(self, index: int):
        self.index = index
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            while not self._stop.is_set():
                self.samples.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                try:
                    r = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(0.1)
        except Exception as e:  # pragma: no cover
            self.reasons.add(f"sampler_error:{type(e).__name__}")

    def start(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": float(s[len(s) // 2]) if s else None, "sm_max_mhz": float(self.max_mhz) if self.max_mhz else None,
                "reasons": sorted(self.reasons), "samples": len(s)}


# --------------------------------------------------------------------------------------------------
# the reference arm / cpu baseline: oracle port on host threads, bounded sample
# --------------------------------------------------------------------------------------------------
def cpu_oracle_rate(protocols, sample: int, threads: int, lo: int = 0):
    """Decode `sample` messages of the mixed corpus (same 40/40/15/5 mix) with the C oracle; returns (msgs, seconds)."""
    from corpus.corpus import Corpus
    from oracle.oracle import Oracle

    corp, ora = Corpus(protocols), Oracle(protocols)
    counts = shard_counts(sample)
    batches = []
    for (name, kind, _), c in zip(MIX, counts):
        if kind <= 1:
            batches.append((kind, corp.pulse(kind, lo + c, lo=lo, hi=lo + c)))
        else:
            batches.append((kind, corp.hexmsgs(kind, lo + c, lo=lo, hi=lo + c)))
    t0 = time.perf_counter()
    for kind, b in batches:
        if kind <= 1:
            ora.run_pulse_raw(b, nthreads=threads)
        else:
            ora.run_hex_raw(b, mc_repaired=True, nthreads=threads)
    return sum(counts), time.perf_counter() - t0


def stratified_dicts(corp, kind: int, total: int, want: int, strata: int = 50, base: int = 0):
    """`want` messages of class `kind` as parser dicts, taken as `strata` evenly spaced runs out of messages
    [base, base + total) of the class's corpus (message i is a pure function of (seed, i)).  Returns (dicts, batches)."""
    from corpus.corpus import batch_to_dicts

    want = min(want, total)
    strata = max(1, min(strata, want))
    per = want // strata
    dicts, batches = [], []
    for s in range(strata):
        lo = base + (total // strata) * s
        hi = lo + per
        b = corp.pulse(kind, base + total, lo=lo, hi=hi) if kind <= 1 else corp.hexmsgs(kind, base + total, lo=lo, hi=hi)
        batches.append(b)
        dicts.extend(batch_to_dicts(b))
    return dicts, batches


def python_reference_rates(protocols, pool, per_class: int, total_per_class, time_cap_s: float = 25.0, keep_results: bool = False):
    """The REAL reference on `pool` (oracle/ref_pool.py): per message class a stratified sample of up to `per_class` messages
    (cut short when the class would take longer than `time_cap_s`).  Returns {name: {...}} and the 40/40/15/5 aggregate."""
    from corpus.corpus import Corpus

    corp = Corpus(protocols)
    out, inv = {}, 0.0
    for (name, kind, share), total in zip(MIX, total_per_class):
        probe, _ = stratified_dicts(corp, kind, total, 64 * pool.workers, strata=8)
        pool.decode(name, probe[: 8 * pool.workers])                       # warm every worker
        _, dt = pool.decode(name, probe)
        rate = len(probe) / dt
        want = int(min(per_class, max(len(probe), rate * time_cap_s)))
        dicts, batches = stratified_dicts(corp, kind, total, want)
        res, dt = pool.decode(name, dicts)
        rate = len(dicts) / dt
        out[name] = {"messages": len(dicts), "seconds": dt, "msgs_per_s": rate, "msgs_per_s_per_core": rate / pool.workers,
                     "hits": sum(len(r[1]) for r in res), "raised": sum(1 for r in res if r[0] != "ok")}
        if keep_results:
            out[name]["_results"] = res
            out[name]["_batches"] = batches
        inv += share / rate
    return out, 1.0 / inv


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from pysignalduino_b200.protocol_data import load_protocol_table

    protocols = load_protocol_table()
    threads = os.cpu_count() or 1
    counts = shard_counts(args.messages)
    from oracle import ref_pool

    if ref_pool.available() and not args.port:
        # the reference's own Python path, Pool(os.cpu_count()); one step = a stratified sample of every class
        from oracle import ref_import

        ref_root = str(ref_import.REFERENCE_ROOT).replace(str(ROOT) + "/", "")
        pool = ref_pool.ReferencePool(threads)
        per_class = max(2000, int(50000 / max(1, args.steps)))            # >= 50 k per class over the run (MU ~ 3 k msg/s on 16 cores)
        for _ in range(min(args.warmup, 1)):
            python_reference_rates(protocols, pool, 32 * threads, counts, time_cap_s=2.0)
        t0 = time.perf_counter()
        agg, per = [], {}
        for k in range(args.steps):
            cls, a = python_reference_rates(protocols, pool, per_class, counts, time_cap_s=max(3.0, 60.0 / max(1, args.steps)))
            agg.append(a)
            for name, v in cls.items():
                p = per.setdefault(name, {"messages": 0, "seconds": 0.0, "hits": 0, "raised": 0})
                for key in ("messages", "seconds", "hits", "raised"):
                    p[key] += v[key]
        pool.close()
        wall = time.perf_counter() - t0
        inv = 0.0
        for (name, _, share) in MIX:
            per[name]["msgs_per_s"] = per[name]["messages"] / per[name]["seconds"]
            per[name]["msgs_per_s_per_core"] = per[name]["msgs_per_s"] / threads
            inv += share / per[name]["msgs_per_s"]
        value = 1.0 / inv
        kind = "reference"
        sample = (f"per step a stratified sample of up to {per_class} messages per class (MS / MU / MC / MN) of the same corpus, decoded by "
                  f"the unmodified Python reference ({ref_root}) in Pool({threads}); value = 40/40/15/5-weighted harmonic rate over {args.steps} steps")
        ms_per_step = 1e3 * wall / max(1, args.steps)
        extra = {"per_class": per}
        note = "reference arm = the reference's own Python SDProtocols.demodulate (MC in the repaired mode of SURVEY 8c), dict inputs prebuilt"
    else:
        n, dt = cpu_oracle_rate(protocols, 20000, threads)                 # calibrate
        rate = n / dt
        per_step = int(min(max(rate * 8.0, 20000), 2_000_000))            # ~8 s of CPU work per step
        for w in range(args.warmup):
            cpu_oracle_rate(protocols, min(per_step, 50000), threads)
        tot_n, tot_t = 0, 0.0
        for k in range(args.steps):
            n, dt = cpu_oracle_rate(protocols, per_step, threads, lo=k * per_step)
            tot_n += n
            tot_t += dt
        value = tot_n / tot_t
        kind = "port"
        sample = f"{per_step} messages/step of the same mixed corpus (40/40/15/5 %), {args.steps} steps, C oracle port on {threads} threads"
        ms_per_step = 1e3 * tot_t / max(1, args.steps)
        extra = {}
        note = "reference arm = C oracle port of the reference's Python path (oracle/_ref absent: run python oracle/make_ref.py in the build container)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": config_dict(args, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample, **extra},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "note": note,
    }
    emit(line)


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
def algorithmic_ops(table, kind_name: str, msgs: np.ndarray, hit_bits: int, n: int):
    """Ops = sum over candidate protocols c of [K_c*8 (tolerance compares) + S_c*dlen (substring / start scans)
    + dlen/w_c (symbol classify) + nbits_c (emit)], K_c = distinct template values, S_c = target strings tested.  MU: every
    protocol is a candidate; MS: the protocols passing the 30 % clock gate (message_synced.py:83-88), evaluated here on a
    sample of the shard.  Returns total thread-level integer ops of one pass over the shard."""
    rows = table.pulse_rows(kind_name)
    K = rows["key"]["nuniq"].astype(np.int64).sum(axis=1)
    S = (rows["key"]["len"] > 0).astype(np.int64).sum(axis=1)
    invw = 1.0 / rows["width"].astype(np.float64)
    dlen = msgs["dlen"].astype(np.float64)
    if kind_name == "MU":
        a, b = float((8 * K).sum()), float((S + invw).sum())
        return a * n + b * float(dlen.sum()) + hit_bits
    sample = msgs[:: max(1, n // 200000)]
    cp = sample["cp"].astype(np.int64)
    ok = cp != 0xFF
    clk = np.abs(sample["pat"][np.arange(len(sample)), np.where(ok, cp, 0)].astype(np.float64))
    pclk = rows["clock"].astype(np.float64)[None, :]
    cand = ok[:, None] & (clk[:, None] != 0) & ~((pclk > 0) & (np.abs(pclk - clk[:, None]) > clk[:, None] * 0.3))
    per_msg = (cand * (8 * K)[None, :]).sum(axis=1) + (cand * (S + invw)[None, :]).sum(axis=1) * sample["dlen"].astype(np.float64)
    return float(per_msg.mean()) * n + hit_bits


def run_ours(args):
    import torch
    import torch.distributed as dist

    from corpus.corpus import Corpus
    from pysignalduino_b200 import SDProtocols, pack
    from pysignalduino_b200.capi import LINEINFO_DTYPE

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the demodulator has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # one process per GPU: run next to this GPU (NUMA) before any pinned buffer is allocated — the e2e path moves ~2 GB per step
    # and rank through host memory
    from pysignalduino_b200.capi import bind_to_gpu_numa

    numa_cpus = None if args.no_numa else bind_to_gpu_numa(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    sdp = SDProtocols(device=local, mc_repaired=True)
    eng = sdp.engine()
    corp = Corpus(sdp.get_protocol_list())
    strong = args.scaling == "strong"
    M = args.messages // world if strong else args.messages          # messages of this rank
    counts = shard_counts(M)
    t_gen = time.perf_counter()
    batches = []
    for (name, kind, _), c in zip(MIX, counts):
        lo = rank * c
        b = corp.pulse(kind, world * c, lo=lo, hi=lo + c) if kind <= 1 else corp.hexmsgs(kind, world * c, lo=lo, hi=lo + c)
        batches.append((name, kind, b))
    t_gen = time.perf_counter() - t_gen

    # ---- device-resident buffers (torch owns the memory and the stream) ----
    def dev_u8(a: np.ndarray):
        return torch.from_numpy(a.view(np.uint8).reshape(-1)).to(dev)

    stream = torch.cuda.current_stream().cuda_stream
    slots = []
    for name, kind, b in batches:
        n = b.n
        hits_cap = max(1024, 8 * n if kind == 1 else 3 * n)
        bits_cap = max(4096, 24 * n if kind == 1 else 8 * n)
        s = {
            "name": name, "kind": kind, "n": n, "batch": b,
            "d_msgs": dev_u8(b.msgs), "d_digits": dev_u8(b.digits),
            "d_out": torch.empty(8 * n, dtype=torch.uint8, device=dev),
            "d_hits": torch.empty(16 * hits_cap, dtype=torch.uint8, device=dev),
            "d_bits": torch.empty(bits_cap, dtype=torch.int32, device=dev),
            "d_ctr": torch.zeros(4, dtype=torch.int32, device=dev),
            "hits_cap": hits_cap, "bits_cap": bits_cap,
        }
        slots.append(s)

    def launch(s, n=None):
        n = s["n"] if n is None else n
        if s["kind"] <= 1:
            eng.demod_pulse_device(s["kind"], s["d_msgs"].data_ptr(), s["d_digits"].data_ptr(), n, s["d_out"].data_ptr(),
                                   s["d_hits"].data_ptr(), s["hits_cap"], s["d_bits"].data_ptr(), s["bits_cap"],
                                   s["d_ctr"].data_ptr(), stream)
        else:
            eng.demod_hex_device(s["kind"], True, s["d_msgs"].data_ptr(), s["d_digits"].data_ptr(), n, s["d_out"].data_ptr(),
                                 s["d_hits"].data_ptr(), s["hits_cap"], s["d_bits"].data_ptr(), s["bits_cap"],
                                 s["d_ctr"].data_ptr(), stream)

    # our kernels per step, per SDB_MU_CHUNK = 1048576 resident messages (sdb_pulse.cu launch_pulse): the counter fold, then
    # MS = resolve + overflow pass + scan, MU = resolve + overflow pass + match + emit + fused fallback, each followed by the two
    # long-message kernels (resolve + scan); the overflow pass and the long kernels exit at once when their list is empty.  One
    # more counter fold ends a call.  MC and MN: one launch each.
    CHUNK = 1048576

    def launches(nmsgs):
        return sum(((6 if s["kind"] == 0 else 8) * ((nm + CHUNK - 1) // CHUNK) + 1) if s["kind"] <= 1 else 1 for s, nm in zip(slots, nmsgs))

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize()

    def allmax(x: float) -> float:
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(args.warmup):
        for s in slots:
            launch(s)
    torch.cuda.synchronize()
    ctrs = {s["name"]: s["d_ctr"].cpu().numpy().astype(np.uint32) for s in slots}
    for s in slots:
        c = ctrs[s["name"]]
        if c[0] > s["hits_cap"] or c[1] > s["bits_cap"]:
            raise SystemExit(f"bench.py: {s['name']} output arena too small: {c}")

    # ---- timed region: K steps, CUDA events on the launching stream ----
    def timed_device(nmsgs):
        ev = [[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in slots] for _ in range(args.steps)]
        e_start, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e_start.record()
        for k in range(args.steps):
            for i, s in enumerate(slots):
                ev[k][i][0].record()
                launch(s, nmsgs[i])
                ev[k][i][1].record()
        e_end.record()
        barrier()
        ms = allmax(e_start.elapsed_time(e_end))
        return ms, [float(np.mean([ev[k][i][0].elapsed_time(ev[k][i][1]) for k in range(args.steps)])) for i in range(len(slots))]

    sampler = ClockSampler(local)
    sampler.start()
    full = [s["n"] for s in slots]
    elapsed_ms, kern_ms = timed_device(full)
    clocks = sampler.stop()
    value = world * M * args.steps / (elapsed_ms / 1e3)

    # ---- end to end through the C ABI with pinned HOST buffers: H2D + kernels + D2H + the payload string of every hit ----
    host = []
    h2d = d2h = 0
    for s in slots:
        b = s["batch"]

        def pin(a: np.ndarray):
            tns = torch.from_numpy(a.view(np.uint8).reshape(-1).copy()).pin_memory()
            return tns, tns.numpy().view(a.dtype).reshape(a.shape)

        t_msgs, msgs = pin(np.ascontiguousarray(b.msgs))
        t_dig, digits = pin(np.ascontiguousarray(b.digits))
        t_out = torch.empty(8 * s["n"], dtype=torch.uint8).pin_memory()
        t_hits = torch.empty(16 * s["hits_cap"], dtype=torch.uint8).pin_memory()
        t_bits = torch.empty(4 * s["bits_cap"], dtype=torch.uint8).pin_memory()
        t_ctr = torch.zeros(16, dtype=torch.uint8).pin_memory()
        c = ctrs[s["name"]]
        t_pool = torch.empty(int(c[0]) * 48 + 4096, dtype=torch.uint8).pin_memory()
        t_off = torch.empty(pack.PAYHIT_DTYPE.itemsize * s["hits_cap"], dtype=torch.uint8).pin_memory()
        host.append({
            "kind": s["kind"], "keep": (t_msgs, t_dig, t_out, t_hits, t_bits, t_ctr, t_pool, t_off), "msgs": msgs, "digits": digits,
            "out": t_out.numpy().view(pack.MSGOUT_DTYPE), "hits": t_hits.numpy().view(pack.HIT_DTYPE),
            "bits": t_bits.numpy().view(np.uint32), "ctr": t_ctr.numpy().view(pack.COUNTERS_DTYPE),
            "pool": t_pool.numpy(), "phits": t_off.numpy().view(pack.PAYHIT_DTYPE), "used": 0,
        })
        h2d += msgs.nbytes + digits.nbytes
        # result slots + counters + 12-byte payload hit records (+ the strings, counted after the first e2e step)
        d2h += 8 * s["n"] + 16 + pack.PAYHIT_DTYPE.itemsize * int(c[0])

    e2e_kind_s = [0.0] * len(slots)
    # One handle (engine) per message kind, driven by its own host thread: the four host calls of a step are in flight together,
    # so the pipeline head (first H2D) and tail (last D2H) of one kind run under the kernels of another and the small MC / MN
    # calls disappear behind MS / MU.  ("one handle may be used by one host thread at a time", include/sdb200.h.)
    from concurrent.futures import ThreadPoolExecutor

    concurrent_kinds = not args.e2e_sequential
    engines = [eng] + [SDProtocols(device=local, mc_repaired=True).engine() for _ in slots[1:]] if concurrent_kinds else [eng] * len(slots)
    pool_exec = ThreadPoolExecutor(max_workers=len(slots)) if concurrent_kinds else None

    def e2e_kind(i, nmsgs):
        hs = host[i]
        t_k = time.perf_counter()
        msgs, digits = hs["msgs"], hs["digits"]
        if nmsgs is not None and nmsgs[i] < len(msgs):
            msgs = msgs[: nmsgs[i]]
            digits = digits[: int(hs["msgs"]["doff"][nmsgs[i]]) * 16 + 64]
        rc, used = engines[i].demod_host_payloads_into(hs["kind"], msgs, digits, hs["out"], hs["phits"], hs["ctr"], hs["pool"],
                                                       mc_repaired=True, bits_cap=len(hs["bits"]))
        if rc != 0:
            raise SystemExit("bench.py: e2e arena / payload pool overflow")
        hs["used"] = used
        e2e_kind_s[i] += time.perf_counter() - t_k

    def e2e_step(nmsgs=None):
        if pool_exec is None:
            for i in range(len(host)):
                e2e_kind(i, nmsgs)
        else:
            list(pool_exec.map(lambda i: e2e_kind(i, nmsgs), range(len(host))))

    def timed_e2e(nmsgs=None):
        e2e_step(nmsgs)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step(nmsgs)
        barrier()
        return allmax(time.perf_counter() - t0)

    t_e2e = timed_e2e()
    e2e_value = world * M * args.steps / t_e2e
    e2e_per_kind_ms = {sl["name"]: 1e3 * e2e_kind_s[i] / (args.steps + 1) for i, sl in enumerate(slots)}
    payload_bytes = int(sum(hs["used"] for hs in host))
    d2h += payload_bytes                                                   # the strings come back from the device instead of the bit arena

    # device-resident and host-buffer runs must agree (same hits, same words)
    for s, hs in zip(slots, host):
        c = ctrs[s["name"]]
        assert int(hs["ctr"]["hits"][0]) == int(c[0]) and int(hs["ctr"]["words"][0]) == int(c[1]), "device / host-path results differ"

    # ---- per-protocol hit histogram of the whole job: every rank's counts, summed (off the timed path) ----
    nproto = len(eng.table.ids)
    hist = np.zeros(nproto, dtype=np.int64)
    for s, hs in zip(slots, host):
        hist += np.bincount(hs["phits"]["proto"][: int(hs["ctr"]["hits"][0])].astype(np.int64), minlength=nproto)[:nproto]
    th = torch.from_numpy(hist).to(dev)
    if world > 1:
        dist.all_reduce(th, op=dist.ReduceOp.SUM)
    hist = th.cpu().numpy()

    # ---- the other scaling mode, same run (N > 1): weak run -> the first M/N messages of every class per rank = ONE
    #      --messages corpus sharded over the ranks (BASELINE config 5); strong run -> nothing more to do ----
    other = None
    if world > 1 and not strong:
        sub = [c // world for c in full]
        ms_sub, _ = timed_device(sub)
        t_sub = timed_e2e(sub)
        tot = sum(sub) * world
        other = {"scaling": "strong", "total_messages": tot, "messages_per_gpu": sum(sub),
                 "value": tot * args.steps / (ms_sub / 1e3), "ms_per_step": ms_sub / args.steps,
                 "e2e": {"value": tot * args.steps / t_sub, "unit": UNIT},
                 "note": "one corpus of --messages sharded by message range over the ranks: each rank decodes the first 1/N of each class of its shard"}

    # ---- line-parser row: firmware TEXT lines of the MS / MU shards through sdb_demod_lines_host
    #      (tokenizer kernel + demodulation kernels, pinned host text in, results out), rank 0 only ----
    lines_info = None
    if rank == 0 and not args.no_lines:
        legs, n_lines, text_bytes, t_lines, ok_lines, lines_hits = [], 0, 0, 0.0, 0, 0
        for s, hs in zip(slots, host):
            if s["kind"] > 1:
                continue
            text, off, ln = corp.render_lines(s["batch"])
            t_text = torch.from_numpy(text.copy()).pin_memory()
            t_info = torch.zeros(s["n"] * LINEINFO_DTYPE.itemsize, dtype=torch.uint8).pin_memory()   # pageable D2H would serialise the stages
            info = t_info.numpy().view(LINEINFO_DTYPE)
            t_off, t_ln = torch.from_numpy(off.copy()).pin_memory(), torch.from_numpy(ln.copy()).pin_memory()
            off, ln = t_off.numpy(), t_ln.numpy()
            legs.append((s, hs, (t_text, t_info, t_off, t_ln), off, ln, info))
            n_lines += s["n"]
            text_bytes += int(text.nbytes)

        def lines_step():
            for s, hs, keep, off, ln, info in legs:
                t_text = keep[0]
                rc = eng.lib.sdb_demod_lines_host(eng.h, s["kind"], t_text.data_ptr(), t_text.numel(), off.ctypes.data, ln.ctypes.data,
                                                  s["n"], hs["out"].ctypes.data, hs["hits"].ctypes.data, len(hs["hits"]),
                                                  hs["bits"].ctypes.data, len(hs["bits"]), hs["ctr"].ctypes.data, info.ctypes.data)
                if rc != 0:
                    raise SystemExit(f"bench.py: sdb_demod_lines_host failed ({rc})")

        if legs:
            lines_step()
            nrep = max(1, min(args.steps, 3))
            t0 = time.perf_counter()
            for _ in range(nrep):
                lines_step()
            t_lines = (time.perf_counter() - t0) / nrep
            for s, hs, t_text, off, ln, info in legs:
                ok_lines += int((info["status"] == 1).sum())
                lines_hits += int(hs["ctr"]["hits"][0])
            # raw receive buffer (STX / ETX framed lines): host framing on all host threads (sdb_frame_lines_inplace), 256 MiB
            # chunks, chunk k + 1 framed while chunk k runs on the device (capi.frame_chunks), same device path as above
            from pysignalduino_b200.capi import frame_chunks, frame_lines_inplace

            t_frame, t_raw, raw_bytes, raw_hits = 0.0, 0.0, 0, 0
            for s, hs, t_text, off, ln, info in legs:
                rtext, _, _ = corp.render_lines(s["batch"], framed=True)
                t_pin = torch.from_numpy(rtext.copy()).pin_memory()          # the receive buffer: pinned, used in place
                raw = memoryview(t_pin.numpy())
                raw_bytes += len(raw)
                t0 = time.perf_counter()
                frame_lines_inplace(raw)                                      # framing alone, for the host-side rate
                t_frame += time.perf_counter() - t0
                t0 = time.perf_counter()
                nl = 0
                for byte_base, byte_end, line_base, foff, fln, ftyp, side in frame_chunks(raw):
                    k = len(ftyp)
                    rc = eng.lib.sdb_demod_lines_host(eng.h, s["kind"], t_pin.data_ptr() + byte_base, byte_end - byte_base,
                                                      foff.ctypes.data, fln.ctypes.data, k, hs["out"].ctypes.data, hs["hits"].ctypes.data,
                                                      len(hs["hits"]), hs["bits"].ctypes.data, len(hs["bits"]), hs["ctr"].ctypes.data,
                                                      info.ctypes.data)
                    if rc != 0 or int((ftyp != s["kind"]).sum()) != 0:
                        raise SystemExit(f"bench.py: raw-buffer leg failed ({rc})")
                    nl += k
                    raw_hits += int(hs["ctr"]["hits"][0])
                t_raw += time.perf_counter() - t0
                if nl != s["n"]:
                    raise SystemExit("bench.py: raw-buffer leg lost lines")
            if raw_hits != lines_hits:
                raise SystemExit(f"bench.py: raw-buffer leg decoded {raw_hits} hits, the payload-line leg {lines_hits}")
            lines_info = {"value": n_lines / t_lines, "unit": "lines/s", "lines": n_lines, "text_bytes_per_step": text_bytes,
                          "decoded_on_device": ok_lines, "hits": lines_hits,
                          "raw_buffer": {"value": n_lines / t_raw, "unit": "lines/s", "bytes": raw_bytes,
                                         "host_framing_lines_per_s": n_lines / t_frame, "host_threads": os.cpu_count(),
                                         "note": "STX/ETX framed receive buffer: sdb_frame_lines_inplace on the host threads in 256 MiB chunks, one chunk ahead of the device path"},
                          "note": "MS + MU shards rendered as firmware payload lines; tokenizer + demodulation kernels, host text in / results out"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ----
    dom = int(np.argmax(kern_ms))
    s = slots[dom]
    b = s["batch"]
    hits_np = s["d_hits"].cpu().numpy().view(pack.HIT_DTYPE)[: int(ctrs[s["name"]][0])]
    if s["kind"] <= 1:
        b_in = 48 * s["n"] + int(((b.msgs["dlen"].astype(np.int64) + 1) // 2).sum())
    else:
        b_in = 16 * s["n"] + int(((b.msgs["hlen"].astype(np.int64) + 1) // 2).sum())
    b_out = int((16 + (hits_np["nbits"].astype(np.int64) + 7) // 8).sum())
    alg_bytes = b_in + b_out
    peaks_path = ROOT / "MEASURED_PEAKS.json"
    if peaks_path.exists():
        peak, peak_src = float(json.loads(peaks_path.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"
    achieved = alg_bytes / (kern_ms[dom] / 1e3) / 1e9
    traffic = None
    issue = None
    tj = None
    tp = ROOT / "profiles" / "traffic.json"
    if tp.exists():
        tj = json.loads(tp.read_text()).get(s["name"])
        if tj and tj.get("messages"):
            traffic = tj["dram_bytes_per_message"] * s["n"]
    if s["kind"] <= 1 and clocks.get("sm_mhz"):
        # the roof that binds: int32 issue, 148 SMs x 4 schedulers x 32 lanes x f_SM thread-ops/s
        ops = algorithmic_ops(eng.table, s["name"], b.msgs, int(hits_np["nbits"].astype(np.int64).sum()), s["n"])
        pk = 148 * 4 * 32 * clocks["sm_mhz"] * 1e6
        issue = {"bound": "int32 issue", "algorithmic_ops_per_message": ops / s["n"], "achieved": ops / (kern_ms[dom] / 1e3), "peak": pk,
                 "unit": "thread-ops/s", "frac": ops / (kern_ms[dom] / 1e3) / pk,
                 "formula": "sum_c [8 K_c + S_c dlen + dlen / w_c + nbits_c] over the candidate protocols (SURVEY 8d)"}
        if tj and tj.get("warp_instructions_per_message"):
            wi = tj["warp_instructions_per_message"] * s["n"] / (kern_ms[dom] / 1e3)
            issue["executed"] = {"warp_instructions_per_message": tj["warp_instructions_per_message"],
                                 "issue_slot_utilisation": wi / (148 * 4 * clocks["sm_mhz"] * 1e6), "source": tj.get("source")}
    roofline = {
        "bound": "issue" if issue else "hbm",
        "kernel": {0: "resolve_kernel<MS> + scan_kernel<MS> (one MS pass)",
                   1: "resolve_kernel<MU> + mu_match_kernel + mu_emit_kernel (one MU pass)"}.get(s["kind"], "hex_kernel"),
        "achieved": achieved,
        "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
        "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": kern_ms[dom], "issue": issue,
        "note": "scan / codec work bound by integer issue, not by HBM: achieved / peak / frac are the HBM figures the contract asks for, "
                "`issue` is the roof that binds (algorithmic ops vs 148 x 4 x 32 x f_SM); see profiles/",
    }
    per_kernel = {sl["name"]: {"messages": sl["n"], "ms": kern_ms[i], "msgs_per_s": sl["n"] / (kern_ms[i] / 1e3),
                               "hits": int(ctrs[sl["name"]][0]), "raised": int(ctrs[sl["name"]][2])} for i, sl in enumerate(slots)}

    # ---- scalar and dict-level API (what a one-line-at-a-time caller pays; signalduino/controller.py:252) ----
    api = None
    if world == 1 and not args.no_cpu:
        from corpus.corpus import batch_to_dicts

        api = {}
        for sl in slots[:2]:
            dicts = batch_to_dicts(corp.pulse(sl["kind"], 20000))
            one = next(d for d in dicts if d.get("data") and len(d["data"]) > 40)
            sdp.demodulate(one, sl["name"])
            t0 = time.perf_counter()
            for _ in range(300):
                sdp.demodulate(one, sl["name"])
            api[f"scalar_latency_us_{sl['name']}"] = (time.perf_counter() - t0) / 300 * 1e6
            sdp.demodulate_batch(dicts[:2000], sl["name"])
            t0 = time.perf_counter()
            sdp.demodulate_batch(dicts, sl["name"])
            api[f"dict_api_msgs_per_s_{sl['name']}"] = len(dicts) / (time.perf_counter() - t0)

    # ---- CPU baselines on the host cores (bounded samples), rank 0 at N = 1 only; their results are the parity check ----
    cpu = None
    parity = None
    if world == 1 and not args.no_cpu:
        from tests.common import compare_raw
        from oracle.oracle import Oracle

        threads = os.cpu_count() or 1
        protocols = sdp.get_protocol_list()
        # (1) C oracle port: timed on the first `sample` messages of every class, then compared hit by hit (protocol, bit
        #     length, payload bytes) with the GPU output of exactly those rows
        n, dt = cpu_oracle_rate(protocols, 20000, threads)
        sample = int(min(max(n / dt * 10.0, 20000), 2_000_000, M))
        ora = Oracle(protocols)
        sc = shard_counts(sample)
        checked_port, t_port = 0, 0.0
        for sl, c in zip(slots, sc):
            sub = corp.pulse(sl["kind"], c) if sl["kind"] <= 1 else corp.hexmsgs(sl["kind"], c)
            t0 = time.perf_counter()
            raw = ora.run_pulse_raw(sub, nthreads=threads) if sl["kind"] <= 1 else ora.run_hex_raw(sub, mc_repaired=True, nthreads=threads)
            t_port += time.perf_counter() - t0
            res = sdp.demodulate_packed(sub)
            why = compare_raw(sdp, sub, res, *raw, check_bits=sl["kind"] <= 1)      # MC / MN oracle hits carry no bit length
            if why:
                raise SystemExit(f"bench.py: PARITY FAILURE against the oracle port, class {sl['name']}: {why}")
            checked_port += c
        port = {"value": sum(sc) / t_port, "unit": UNIT, "cores": threads, "kind": "port",
                "sample": f"first {sample} messages of the same mixed corpus (40/40/15/5 %), C oracle port on {threads} threads"}
        # (2) the REAL Python reference (oracle/_ref) in Pool(cores) on a stratified sample of every class, compared message by
        #     message with the GPU output of the same rows
        cpu = port
        checked_ref = 0
        from oracle import ref_pool

        if ref_pool.available():
            from oracle import ref_import

            ref_root = str(ref_import.REFERENCE_ROOT).replace(str(ROOT) + "/", "")
            pool = ref_pool.ReferencePool(threads)
            per_class = int(args.ref_sample)
            cls, agg = python_reference_rates(protocols, pool, per_class, counts, time_cap_s=25.0, keep_results=True)
            pool.close()
            for sl in slots:
                exp = cls[sl["name"]].pop("_results")
                k = 0
                for sub in cls[sl["name"]].pop("_batches"):
                    st, rs = sdp.format_results(sub, sdp.demodulate_packed(sub))
                    for a, lst in zip(st, rs):
                        got = (a, [(str(x["protocol_id"]), x["payload"], int(x["meta"].get("bit_length", -1))) for x in lst])
                        want = exp[k]
                        if sl["kind"] >= 2:                      # MC / MN dicts carry no bit_length in meta
                            got = (got[0], [(p_, q_) for p_, q_, _ in got[1]])
                            want = (want[0], [(p_, q_) for p_, q_, _ in want[1]])
                        if got != want:
                            raise SystemExit(f"bench.py: PARITY FAILURE against the Python reference, class {sl['name']} sample row {k}: gpu={got} reference={want}")
                        k += 1
                checked_ref += k
            cpu = {"value": agg, "unit": UNIT, "cores": threads, "kind": "reference",
                   "sample": f"stratified sample of up to {per_class} messages per class (50 runs spread over the shard), unmodified Python reference "
                             f"({ref_root}, MC repaired per SURVEY 8c) in multiprocessing.Pool({threads}); value = 40/40/15/5-weighted harmonic rate",
                   "per_class": cls, "port": port}
        parity = {"parity_checked_messages": checked_port + checked_ref, "against_python_reference": checked_ref,
                  "against_oracle_port": checked_port, "mismatches": 0}

    top = np.argsort(-hist)[:12]
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "int32", "data": "synthetic", "config": config_dict(args, world),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "payload_bytes_per_step": payload_bytes, "host_call_ms_per_kind": e2e_per_kind_ms,
                "host_threads": len(slots) if concurrent_kinds else 1,
                "includes": "pinned H2D, decode kernels, the payload string of every hit (preamble + hex + postamble) by the device format kernel "
                            "of each pipeline stage, D2H of result slots / 12-byte payload hit records / strings"},
        "gpu_launches": args.steps * launches(full),
        "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "parity": parity, "per_kernel": per_kernel,
        "hit_histogram": {"protocols_with_hits": int((hist > 0).sum()), "total_hits": int(hist.sum()), "ranks": world,
                          "top": {eng.table.ids[int(i)]: int(hist[int(i)]) for i in top}},
        "api": api, "other_scaling": other,
        "scratch": dict(flagged_messages=eng.scratch_short(), **eng.scratch_info(),
                        note="compact survivor / match arenas of the device-resident handle (bytes per launch group of `chunk` messages; "
                             "round 1: chunk x 129 x 16 B + chunk x 64 x 4 B = 2.2 GB per 1 048 576 messages)"),
        "host": {"cpus": os.cpu_count(), "rank0_affinity_cpus": len(numa_cpus) if numa_cpus else None,
                 "numa_binding": "rank pinned to its GPU's NVML CPU affinity before pinned buffers are allocated" if numa_cpus else "none"},
        "corpus_gen_s": t_gen, "lines": lines_info,
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--messages", type=int, default=10_000_000, help="messages per GPU (mixed corpus)")
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--scaling", choices=["weak", "strong"], default="weak",
                    help="weak: --messages per GPU (default); strong: ONE corpus of --messages sharded over the GPUs (BASELINE config 5)")
    ap.add_argument("--port", action="store_true", help="--impl reference: time the C oracle port instead of the Python reference")
    ap.add_argument("--ref-sample", type=int, default=50000, help="cpu_baseline: messages per class for the Python reference")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline / parity / API legs")
    ap.add_argument("--e2e-sequential", action="store_true", help="e2e: issue the four kinds one after the other from one thread")
    ap.add_argument("--no-numa", action="store_true", help="do not pin the rank to its GPU's NUMA-local cores")
    ap.add_argument("--no-lines", action="store_true", help="skip the text-line (tokenizer) leg")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    capture_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
