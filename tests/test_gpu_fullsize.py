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
