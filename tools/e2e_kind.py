#!/usr/bin/env python3
"""One message kind alone: device-resident pass vs the host-buffer payload call (pinned buffers), to see what the pipeline of
sdb_demod_host_payloads costs on top of the kernels.   python tools/e2e_kind.py MU 4000000 [reps]"""
import sys, time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch

from corpus.corpus import Corpus
from pysignalduino_b200 import SDProtocols, pack

name = sys.argv[1] if len(sys.argv) > 1 else "MU"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4_000_000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
kind = pack.KIND_BY_NAME[name]
sdp = SDProtocols(device=0, mc_repaired=True)
eng = sdp.engine()
b = Corpus(sdp.get_protocol_list()).pulse(kind, n)
dev = torch.device("cuda", 0)
u8 = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).to(dev)
d_msgs, d_dig = u8(b.msgs), u8(b.digits)
hc, bc = 8 * n, 24 * n
d_out = torch.empty(8 * n, dtype=torch.uint8, device=dev)
d_hits = torch.empty(16 * hc, dtype=torch.uint8, device=dev)
d_bits = torch.empty(bc, dtype=torch.int32, device=dev)
d_ctr = torch.zeros(4, dtype=torch.int32, device=dev)
st = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    eng.demod_pulse_device(kind, d_msgs.data_ptr(), d_dig.data_ptr(), n, d_out.data_ptr(), d_hits.data_ptr(), hc, d_bits.data_ptr(), bc, d_ctr.data_ptr(), st)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    eng.demod_pulse_device(kind, d_msgs.data_ptr(), d_dig.data_ptr(), n, d_out.data_ptr(), d_hits.data_ptr(), hc, d_bits.data_ptr(), bc, d_ctr.data_ptr(), st)
e1.record(); torch.cuda.synchronize()
dev_ms = e0.elapsed_time(e1) / reps
nh = int(d_ctr.cpu().numpy().astype(np.uint32)[0])
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1).copy()).pin_memory()
t_msgs, t_dig = pin(b.msgs), pin(b.digits)
msgs, digits = t_msgs.numpy().view(b.msgs.dtype), t_dig.numpy()
t_out = torch.empty(8 * n, dtype=torch.uint8).pin_memory()
t_ph = torch.empty(pack.PAYHIT_DTYPE.itemsize * hc, dtype=torch.uint8).pin_memory()
t_pool = torch.empty(nh * 48 + 4096, dtype=torch.uint8).pin_memory()
t_ctr = torch.zeros(16, dtype=torch.uint8).pin_memory()
call = lambda: eng.demod_host_payloads_into(kind, msgs, digits, t_out.numpy().view(pack.MSGOUT_DTYPE), t_ph.numpy().view(pack.PAYHIT_DTYPE),
                                            t_ctr.numpy().view(pack.COUNTERS_DTYPE), t_pool.numpy(), mc_repaired=True, bits_cap=bc)
call(); call()
t0 = time.perf_counter()
for _ in range(reps):
    rc, used = call()
host_ms = (time.perf_counter() - t0) / reps * 1e3
print(f"{name} n={n}: device-resident {dev_ms:.2f} ms ({n / dev_ms / 1e3:.1f} M/s), host-buffer payload call {host_ms:.2f} ms ({n / host_ms / 1e3:.1f} M/s), "
      f"rc={rc}, hits={nh}, h2d={(msgs.nbytes + digits.nbytes) / 1e6:.0f} MB, d2h={(8 * n + 12 * nh + used) / 1e6:.0f} MB")
