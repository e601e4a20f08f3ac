"""GPU parity at BASELINE.json sizes: 1 M messages per class, every hit compared (status, protocol,
bit_length, payload bytes) against the C oracle with vectorised array compares."""
import pytest

from pysignalduino_b200 import pack
from tests.common import compare_raw

pytestmark = pytest.mark.gpu
N = 1_000_000


@pytest.mark.parametrize("kind", [pack.KIND_MS, pack.KIND_MU])
def test_one_million_pulse_messages(sdp, oracle, corpus, kind):
    batch = corpus.pulse(kind, N)
    res = sdp.demodulate_packed(batch)
    status, hits, pool = oracle.run_pulse_raw(batch, nthreads=16)
    assert compare_raw(sdp, batch, res, status, hits, pool) == ""


@pytest.mark.parametrize("kind,repaired", [(pack.KIND_MC, True), (pack.KIND_MC, False), (pack.KIND_MN, True)])
def test_one_million_hex_messages(sdp, oracle, corpus, kind, repaired):
    batch = corpus.hexmsgs(kind, N)
    res = sdp.engine().demod_host(batch, mc_repaired=repaired)
    status, hits, pool = oracle.run_hex_raw(batch, mc_repaired=repaired, nthreads=16)
    assert compare_raw(sdp, batch, res, status, hits, pool, check_bits=False) == ""


def test_device_resident_call_crosses_its_launch_chunk(sdp, corpus):
    """sdb_demod_pulse_device on 1.2 M resident MU messages (> SDB_MU_CHUNK = 1 048 576: two launch groups) gives the
    same per-message results as the pipelined host-buffer call (262 144-message stages)."""
    import numpy as np
    import torch

    n = 1_200_000
    batch = corpus.pulse(pack.KIND_MU, n)
    eng = sdp.engine()
    ref = eng.demod_host(batch)
    dev = torch.device("cuda", 0)
    u8 = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).to(dev)  # noqa: E731
    d_msgs, d_dig = u8(batch.msgs), u8(batch.digits)
    hc, bc = len(ref.hits) + 16, len(ref.bits) + 16
    d_out = torch.empty(8 * n, dtype=torch.uint8, device=dev)
    d_hits = torch.empty(16 * hc, dtype=torch.uint8, device=dev)
    d_bits = torch.empty(bc, dtype=torch.int32, device=dev)
    d_ctr = torch.zeros(4, dtype=torch.int32, device=dev)
    eng.demod_pulse_device(pack.KIND_MU, d_msgs.data_ptr(), d_dig.data_ptr(), n, d_out.data_ptr(), d_hits.data_ptr(), hc,
                           d_bits.data_ptr(), bc, d_ctr.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    ctr = d_ctr.cpu().numpy().astype(np.uint32)
    assert int(ctr[0]) == len(ref.hits) and int(ctr[1]) == len(ref.bits) and int(ctr[2]) == int(ref.counters["raised"])
    out = d_out.cpu().numpy().view(pack.MSGOUT_DTYPE)
    assert np.array_equal(out["status"], ref.out["status"]) and np.array_equal(out["nhits"], ref.out["nhits"])
    hits = d_hits.cpu().numpy().view(pack.HIT_DTYPE)[: int(ctr[0])]
    o1 = np.lexsort((np.arange(len(hits)), hits["msg"].astype(np.int64)))
    o2 = np.lexsort((np.arange(len(ref.hits)), ref.hits["msg"].astype(np.int64)))
    for f in ("msg", "proto", "nbits", "aux", "flags"):
        assert np.array_equal(hits[f][o1], ref.hits[f][o2]), f
