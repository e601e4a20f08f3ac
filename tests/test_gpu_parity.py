"""GPU parity: CUDA path (through the C ABI) vs the CPU oracle on the same seeded inputs."""
import numpy as np
import pytest

from pysignalduino_b200 import pack
from tests.common import canonical_gpu, diff_report

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind,n", [(pack.KIND_MS, 20000), (pack.KIND_MU, 20000)])
def test_corpus_parity(sdp, oracle, corpus, kind, n):
    batch = corpus.pulse(kind, n)
    res = sdp.demodulate_packed(batch)
    got = canonical_gpu(sdp, batch, res)
    exp = oracle.run_pulse(batch, nthreads=8)
    assert got == exp, diff_report(got, exp)
    assert int(res.counters["hits"]) == sum(len(e[1]) for e in exp)


@pytest.mark.parametrize("kind", [pack.KIND_MS, pack.KIND_MU])
def test_shard_invariance(sdp, corpus, kind):
    """Message i is a pure function of (seed, i): a shard decodes exactly like the same rows of the whole."""
    whole = corpus.pulse(kind, 4000)
    part = corpus.pulse(kind, 4000, lo=1000, hi=3000)
    a = canonical_gpu(sdp, whole, sdp.demodulate_packed(whole))[1000:3000]
    b = canonical_gpu(sdp, part, sdp.demodulate_packed(part))
    assert a == b


def test_empty_and_invalid(sdp):
    st, res = sdp.demodulate_batch([], "MS")
    assert st == [] and res == []
    msgs = [{"data": ""}, {"data": "0101", "CP": "x", "SP": "0", "P0": "100"}, {"P0": "1", "CP": "0", "SP": "0"}]
    st, res = sdp.demodulate_batch(msgs, "MS")
    assert st == ["ok"] * 3 and res == [[], [], []]
    st, res = sdp.demodulate_batch([{"data": ""}, {"P0": "5"}], "MU")
    assert st == ["ok"] * 2 and res == [[], []]


@pytest.mark.parametrize("kind,repaired", [(pack.KIND_MC, True), (pack.KIND_MC, False), (pack.KIND_MN, True)])
def test_hex_corpus_parity(sdp, oracle, corpus, kind, repaired):
    batch = corpus.hexmsgs(kind, 30000)
    res = sdp.engine().demod_host(batch, mc_repaired=repaired)
    got = canonical_gpu(sdp, batch, res)
    exp = oracle.run_hex(batch, mc_repaired=repaired, nthreads=8)
    assert got == exp, diff_report(got, exp)


@pytest.mark.parametrize("kind,n", [(pack.KIND_MS, 600000), (pack.KIND_MU, 300000)])
def test_pipelined_host_path_parity(sdp, oracle, corpus, kind, n):
    """n > one chunk (262144): the host-buffer call overlaps H2D / kernels / D2H per chunk; MU also crosses
    a resolve/scan chunk boundary.  Same results, hits contiguous per message and in reference order."""
    batch = corpus.pulse(kind, n)
    res = sdp.demodulate_packed(batch)
    status, hits, pool = oracle.run_pulse_raw(batch, nthreads=16)
    assert int(res.counters["hits"]) == len(hits)
    assert np.array_equal(res.out["status"], status)
    # per-message hit lists: (proto, nbits) sequences must agree everywhere ...
    order = np.argsort(res.out["hit_off"], kind="stable")
    g_msg = res.hits["msg"].astype(np.int64)
    assert np.array_equal(np.sort(g_msg, kind="stable"), hits["msg"].astype(np.int64))
    o = np.lexsort((np.arange(len(g_msg)), g_msg))        # stable by message: device hits of one message are contiguous + ordered
    assert np.array_equal(res.hits["proto"][o].astype(np.int64), hits["proto"].astype(np.int64))
    assert np.array_equal(res.hits["nbits"][o].astype(np.int64), hits["bit_length"].astype(np.int64))
    # ... and payload strings on a slice that crosses the chunk boundary
    lo, hi = 262144 - 500, 262144 + 500
    sub = corpus.pulse(kind, n, lo=lo, hi=hi)
    got = canonical_gpu(sdp, sub, sdp.demodulate_packed(sub))
    exp = oracle.run_pulse(sub)
    assert got == exp, diff_report(got, exp)
    statuses, results = sdp.format_results(batch, res)
    whole = [(st, [(r["protocol_id"], r["payload"], r["meta"]["bit_length"]) for r in lst])
             for st, lst in zip(statuses[lo:hi], results[lo:hi])]
    assert whole == exp
