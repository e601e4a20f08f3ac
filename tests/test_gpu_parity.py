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
