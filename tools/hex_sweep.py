"""MC / MN kernel rate for the grid size given by SDB_HEX_CTAS (CTAs per SM):  SDB_HEX_CTAS=48 python tools/hex_sweep.py"""
import sys, os
sys.path.insert(0, ".")
import numpy as np, torch
from corpus.corpus import Corpus
from pysignalduino_b200 import SDProtocols, pack
sdp = SDProtocols(device=0, mc_repaired=True); eng = sdp.engine(); corp = Corpus(sdp.get_protocol_list())
dev = torch.device("cuda", 0)
for name, n in (("MC", 1500000), ("MN", 500000)):
    kind = pack.KIND_BY_NAME[name]; b = corp.hexmsgs(kind, n)
    u8 = lambda a: torch.from_numpy(a.view(np.uint8).reshape(-1)).to(dev)
    dm, dd = u8(b.msgs), u8(b.digits); hc, bc = 4 * n, 16 * n
    do = torch.empty(8 * n, dtype=torch.uint8, device=dev); dh = torch.empty(16 * hc, dtype=torch.uint8, device=dev)
    db = torch.empty(bc, dtype=torch.int32, device=dev); dc = torch.zeros(4, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    f = lambda: eng.demod_hex_device(kind, True, dm.data_ptr(), dd.data_ptr(), n, do.data_ptr(), dh.data_ptr(), hc, db.data_ptr(), bc, dc.data_ptr(), st)
    for _ in range(3): f()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize()
    print("ctas/sm", os.environ.get("SDB_HEX_CTAS"), name, round(n / (e0.elapsed_time(e1) / 10) / 1e3, 1), "M msg/s")
