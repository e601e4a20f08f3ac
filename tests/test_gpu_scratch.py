"""The compact scratch arenas (csrc/sdb_pulse.h: survivor / match records in blocks claimed per warp, sized by AVERAGE counts;
``slack_warps=1`` removes the one-block-per-resident-warp allowance so that small batches exhaust them):
every overflow path must give the results of the roomy default — overflow pass of the resolve kernel, the fused fallback
kernel for a full match arena, SDB_ST_SCRATCH + growth + repetition in the host-buffer calls, and the device-pointer protocol
(flagged messages, sdb_scratch_short, resubmission)."""
import numpy as np
import pytest

from pysignalduino_b200 import SDProtocols, pack
from pysignalduino_b200.capi import ST_SCRATCH
from tests.common import canonical_gpu, diff_report

pytestmark = pytest.mark.gpu


@pytest.fixture()
def own():
    s = SDProtocols()            # its own handle: the budgets are per handle
    yield s
    s.engine().close()


@pytest.mark.parametrize("kind", [pack.KIND_MS, pack.KIND_MU])
def test_overflow_pass_gives_the_same_results(own, oracle, corpus, kind):
    batch = corpus.pulse(kind, 6000)
    exp = oracle.run_pulse(batch, nthreads=8)
    eng = own.engine()
    eng.scratch_budget(surv_avg=1, match_avg=12, ovf_max=8192, slack_warps=1)   # nearly every message with survivors takes the overflow pass
    got = canonical_gpu(own, batch, own.demodulate_packed(batch))
    assert got == exp, diff_report(got, exp)
    info = eng.scratch_info()
    assert info["surv_avg"] == 1 and info["bytes"] > 0


def test_full_match_arena_falls_back_to_the_fused_kernel(own, oracle, corpus):
    batch = corpus.pulse(pack.KIND_MU, 6000)
    exp = oracle.run_pulse(batch, nthreads=8)
    own.engine().scratch_budget(surv_avg=18, match_avg=1, ovf_max=8192, slack_warps=1)
    got = canonical_gpu(own, batch, own.demodulate_packed(batch))
    assert got == exp, diff_report(got, exp)


@pytest.mark.parametrize("kind", [pack.KIND_MS, pack.KIND_MU])
def test_host_calls_grow_the_scratch_and_repeat(own, oracle, corpus, kind):
    """surv_avg 1 and room for 16 overflow messages: most messages are flagged SDB_ST_SCRATCH on the first attempt; the host
    call grows the scratch from the recorded need and repeats, the caller sees complete results."""
    batch = corpus.pulse(kind, 6000)
    exp = oracle.run_pulse(batch, nthreads=8)
    eng = own.engine()
    eng.scratch_budget(surv_avg=1, match_avg=12, ovf_max=16, slack_warps=1)
    res = own.demodulate_packed(batch)
    assert not (res.out["status"] == ST_SCRATCH).any()
    got = canonical_gpu(own, batch, res)
    assert got == exp, diff_report(got, exp)
    info = eng.scratch_info()
    assert info["surv_avg"] > 1 or info["ovf_max"] > 16
    # the payload call (device format kernel) goes through the same repetition
    eng.scratch_budget(surv_avg=1, match_avg=12, ovf_max=16, slack_warps=1)
    res2, pool = eng.demod_payloads(batch)
    assert not (res2.out["status"] == ST_SCRATCH).any()
    assert int(res2.counters["hits"]) == sum(len(e[1]) for e in exp)


def test_device_calls_flag_messages_and_resubmission_succeeds(own, oracle, corpus):
    torch = pytest.importorskip("torch")
    n = 4000
    batch = corpus.pulse(pack.KIND_MU, n)
    exp = oracle.run_pulse(batch, nthreads=8)
    eng = own.engine()
    eng.scratch_budget(surv_avg=1, match_avg=12, ovf_max=16, slack_warps=1)
    dev = torch.device("cuda", 0)
    u8 = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).to(dev)
    d_msgs, d_dig = u8(batch.msgs), u8(batch.digits)
    hc, bc = 32 * n, 64 * n
    d_out = torch.zeros(8 * n, dtype=torch.uint8, device=dev)
    d_hits = torch.zeros(16 * hc, dtype=torch.uint8, device=dev)
    d_bits = torch.zeros(bc, dtype=torch.int32, device=dev)
    d_ctr = torch.zeros(4, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream().cuda_stream

    def run():
        eng.demod_pulse_device(pack.KIND_MU, d_msgs.data_ptr(), d_dig.data_ptr(), n, d_out.data_ptr(), d_hits.data_ptr(), hc,
                               d_bits.data_ptr(), bc, d_ctr.data_ptr(), st)
        torch.cuda.synchronize()
        return d_out.cpu().numpy().view(np.dtype([("hit_off", "<u4"), ("nhits", "<u2"), ("status", "u1"), ("reason", "u1")]))

    out = run()
    flagged = int((out["status"] == ST_SCRATCH).sum())
    assert flagged > 0
    assert eng.scratch_short() == flagged          # ... and the budgets grew
    assert eng.scratch_short() == 0
    # a message that was NOT flagged is complete and correct already
    done = np.flatnonzero(out["status"] != ST_SCRATCH)
    assert [int(out["nhits"][i]) for i in done[:500]] == [len(exp[i][1]) for i in done[:500]]
    out2 = run()                                   # the same batch again: now everything fits
    assert not (out2["status"] == ST_SCRATCH).any()
    assert eng.scratch_short() == 0
    assert [int(x) for x in out2["nhits"]] == [len(e[1]) for e in exp]


def test_lines_call_grows_the_scratch_and_repeats(own, corpus):
    """sdb_demod_lines_host (tokenizer + decode kernels) goes through the same growth and repetition: with budgets that flag
    most messages on the first attempt its results equal those of the roomy default."""
    kind = pack.KIND_MU
    n = 5000
    b = corpus.pulse(kind, n)
    lines = []
    for i in range(n):
        d = pack.unpack_pulse(b, i)
        if not d.get("data"):
            d = {"data": "", "P0": "1"}
        lines.append((";".join(["MU"] + [f"{k}={v}" for k, v in d.items() if k.startswith("P")] + [f"D={d['data']}"]) + ";").encode("ascii"))
    lens = np.fromiter((len(x) for x in lines), dtype=np.int64, count=n)
    offs = np.zeros(n, dtype=np.int64)
    np.cumsum(lens[:-1] + 1, out=offs[1:])
    text = np.frombuffer(b"\n".join(lines) + b"\n", dtype=np.uint8)
    eng = own.engine()
    ref, ref_info = eng.demod_lines(kind, text, offs.astype(np.uint32), lens.astype(np.uint32))
    eng.scratch_budget(surv_avg=1, match_avg=8, ovf_max=16, slack_warps=1)
    res, info = eng.demod_lines(kind, text, offs.astype(np.uint32), lens.astype(np.uint32))
    assert not (res.out["status"] == ST_SCRATCH).any()
    assert np.array_equal(info["status"], ref_info["status"])
    assert np.array_equal(res.out["status"], ref.out["status"]) and np.array_equal(res.out["nhits"], ref.out["nhits"])
    assert int(res.counters["hits"]) == int(ref.counters["hits"]) > 0
    o1 = np.lexsort((np.arange(len(res.hits)), res.hits["msg"].astype(np.int64)))
    o2 = np.lexsort((np.arange(len(ref.hits)), ref.hits["msg"].astype(np.int64)))
    for f in ("msg", "proto", "nbits", "aux", "flags"):
        assert np.array_equal(res.hits[f][o1], ref.hits[f][o2]), f
