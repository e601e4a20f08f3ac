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
           (pinned H2D + kernels + D2H inside the timed region); `roofline` for the dominant kernel
           (MU) with algorithmic bytes 48 + ceil(dlen/2) + sum_hits(16 + ceil(nbits/8)) per message.
reference: the CPU oracle port (oracle/, C restatement of the reference's Python path, pinned
           against the reference here) on all host threads, on a bounded sample of the same corpus.
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
    return {
        "workload": "config5: mixed MS/MU/MC/MN corpus (40/40/15/5 %), seed 0x5D05 family, x all protocols of each class",
        "messages_per_gpu": args.messages,
        "total_messages": args.messages * n_gpus,
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


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from pysignalduino_b200.protocol_data import load_protocol_table

    protocols = load_protocol_table()
    threads = os.cpu_count() or 1
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
    sample = f"{per_step} messages/step of the same mixed corpus (40/40/15/5 %), {args.steps} steps"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / max(1, args.steps), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": config_dict(args, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference arm = C oracle port of the reference's Python path (the Python reference cannot travel to the GPU box); "
                "survey-time probe of the real Python reference: ~1.7k MS msg/s/core, ~450 MU msg/s/core",
    }
    emit(line)


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
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
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    sdp = SDProtocols(device=local, mc_repaired=True)
    eng = sdp.engine()
    corp = Corpus(sdp.get_protocol_list())
    M = args.messages
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

    def launch(s):
        if s["kind"] <= 1:
            eng.demod_pulse_device(s["kind"], s["d_msgs"].data_ptr(), s["d_digits"].data_ptr(), s["n"], s["d_out"].data_ptr(),
                                   s["d_hits"].data_ptr(), s["hits_cap"], s["d_bits"].data_ptr(), s["bits_cap"],
                                   s["d_ctr"].data_ptr(), stream)
        else:
            eng.demod_hex_device(s["kind"], True, s["d_msgs"].data_ptr(), s["d_digits"].data_ptr(), s["n"], s["d_out"].data_ptr(),
                                 s["d_hits"].data_ptr(), s["hits_cap"], s["d_bits"].data_ptr(), s["bits_cap"],
                                 s["d_ctr"].data_ptr(), stream)

    # our kernels per step: MS and MU per SDB_MU_CHUNK = 1048576 resident messages, MC 1, MN 1
    CHUNK = 1048576
    # per chunk: MS = resolve + scan, MU = resolve + match + emit + fused fallback (sdb_pulse.cu launch_pulse)
    launches_per_step = sum(((2 if s["kind"] == 0 else 4) * ((s["n"] + CHUNK - 1) // CHUNK)) if s["kind"] <= 1 else 1 for s in slots)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize()

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
    sampler = ClockSampler(local)
    sampler.start()
    ev = [[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in slots] for _ in range(args.steps)]
    e_start, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e_start.record()
    for k in range(args.steps):
        for i, s in enumerate(slots):
            ev[k][i][0].record()
            launch(s)
            ev[k][i][1].record()
    e_end.record()
    barrier()
    clocks = sampler.stop()
    elapsed_ms = e_start.elapsed_time(e_end)
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    value = world * M * args.steps / (elapsed_ms / 1e3)
    kern_ms = [float(np.mean([ev[k][i][0].elapsed_time(ev[k][i][1]) for k in range(args.steps)])) for i in range(len(slots))]

    # ---- end to end through the C ABI with pinned HOST buffers ----
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
        host.append({
            "kind": s["kind"], "keep": (t_msgs, t_dig, t_out, t_hits, t_bits, t_ctr), "msgs": msgs, "digits": digits,
            "out": t_out.numpy().view(pack.MSGOUT_DTYPE), "hits": t_hits.numpy().view(pack.HIT_DTYPE),
            "bits": t_bits.numpy().view(np.uint32), "ctr": t_ctr.numpy().view(pack.COUNTERS_DTYPE),
        })
        c = ctrs[s["name"]]
        h2d += msgs.nbytes + digits.nbytes
        d2h += 8 * s["n"] + 16 + 16 * int(c[0]) + 4 * int(c[1])

    def e2e_step():
        for hs in host:
            rc = eng.demod_host_into(hs["kind"], hs["msgs"], hs["digits"], hs["out"], hs["hits"], hs["bits"], hs["ctr"], mc_repaired=True)
            if rc != 0:
                raise SystemExit("bench.py: e2e arena overflow")

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    t_e2e = time.perf_counter() - t0
    t = torch.tensor([t_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * M * args.steps / float(t.item())

    # device-resident and host-buffer runs must agree (same hits, same words)
    for s, hs in zip(slots, host):
        c = ctrs[s["name"]]
        assert int(hs["ctr"]["hits"][0]) == int(c[0]) and int(hs["ctr"]["words"][0]) == int(c[1]), "device / host-path results differ"

    # ---- line-parser row: firmware TEXT lines of the MS / MU shards through sdb_demod_lines_host
    #      (tokenizer kernel + demodulation kernels, pinned host text in, results out), rank 0 only ----
    lines_info = None
    if rank == 0 and not args.no_lines:
        import ctypes as C

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
    tp = ROOT / "profiles" / "traffic.json"
    if tp.exists():
        tj = json.loads(tp.read_text()).get(s["name"])
        if tj and tj.get("messages"):
            traffic = tj["dram_bytes_per_message"] * s["n"]
            if tj.get("warp_instructions_per_message") and clocks.get("sm_mhz"):
                # second roofline, the one that binds: warp-instruction issue slots (148 SMs x 4 schedulers x SM clock)
                wi = tj["warp_instructions_per_message"] * s["n"] / (kern_ms[dom] / 1e3)
                pk = 148 * 4 * clocks["sm_mhz"] * 1e6
                issue = {"warp_instructions_per_message": tj["warp_instructions_per_message"], "achieved": wi, "peak": pk,
                         "unit": "warp-instructions/s", "frac": wi / pk, "source": tj.get("source")}
    roofline = {
        "bound": "hbm", "kernel": {0: "resolve_kernel<MS> + scan_kernel<MS> (one MS pass)",
                                   1: "resolve_kernel<MU> + mu_match_kernel + mu_emit_kernel (one MU pass)"}.get(s["kind"], "hex_kernel"),
        "achieved": achieved,
        "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
        "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": kern_ms[dom], "issue": issue,
        "note": "integer-issue bound scan/codec work: see profiles/ for issue-slot utilisation; HBM fraction is low by construction",
    }
    per_kernel = {sl["name"]: {"messages": sl["n"], "ms": kern_ms[i], "msgs_per_s": sl["n"] / (kern_ms[i] / 1e3),
                               "hits": int(ctrs[sl["name"]][0]), "raised": int(ctrs[sl["name"]][2])} for i, sl in enumerate(slots)}

    # ---- CPU baseline on the host cores (bounded sample), rank 0 at N = 1 only ----
    cpu = None
    if world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        n, dt = cpu_oracle_rate(sdp.get_protocol_list(), 20000, threads)
        sample = int(min(max(n / dt * 10.0, 20000), 2_000_000))
        n, dt = cpu_oracle_rate(sdp.get_protocol_list(), sample, threads)
        cpu = {"value": n / dt, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"first {sample} messages of the same mixed corpus (40/40/15/5 %), C oracle port on {threads} threads"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int32", "data": "synthetic", "config": config_dict(args, world),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
        "gpu_launches": args.steps * launches_per_step,
        "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "per_kernel": per_kernel,
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
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
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
