#!/usr/bin/env python3
"""Small device-resident run of one kernel for ncu (`--set full`): tools/profile_run.py MU 40000 3"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch

from corpus.corpus import Corpus
from pysignalduino_b200 import SDProtocols, pack

name = sys.argv[1] if len(sys.argv) > 1 else "MU"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
kind = pack.KIND_BY_NAME[name]
sdp = SDProtocols(device=0, mc_repaired=True)
eng = sdp.engine()
corp = Corpus(sdp.get_protocol_list())
b = corp.pulse(kind, n) if kind <= 1 else corp.hexmsgs(kind, n)
dev = torch.device("cuda", 0)
u8 = lambda a: torch.from_numpy(a.view(np.uint8).reshape(-1)).to(dev)
d_msgs, d_dig = u8(b.msgs), u8(b.digits)
hc, bc = 16 * n, 48 * n
d_out = torch.empty(8 * n, dtype=torch.uint8, device=dev)
d_hits = torch.empty(16 * hc, dtype=torch.uint8, device=dev)
d_bits = torch.empty(bc, dtype=torch.int32, device=dev)
d_ctr = torch.zeros(4, dtype=torch.int32, device=dev)
st = torch.cuda.current_stream().cuda_stream
for _ in range(reps):
    if kind <= 1:
        eng.demod_pulse_device(kind, d_msgs.data_ptr(), d_dig.data_ptr(), n, d_out.data_ptr(), d_hits.data_ptr(), hc,
                               d_bits.data_ptr(), bc, d_ctr.data_ptr(), st)
    else:
        eng.demod_hex_device(kind, True, d_msgs.data_ptr(), d_dig.data_ptr(), n, d_out.data_ptr(), d_hits.data_ptr(), hc,
                             d_bits.data_ptr(), bc, d_ctr.data_ptr(), st)
torch.cuda.synchronize()
print(name, n, d_ctr.cpu().numpy())
