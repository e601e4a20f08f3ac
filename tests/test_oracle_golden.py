"""CPU: the C oracle (and the host packer) against reference-generated golden fixtures."""
import pytest

from pysignalduino_b200 import pack
from tests.common import diff_report, golden_expected, load_golden

PULSE = [("reference_vectors.json.gz", None), ("corpus_ms.json.gz", "MS"), ("corpus_mu.json.gz", "MU"),
         ("fuzz_ms.json.gz", "MS"), ("fuzz_mu.json.gz", "MU"), ("crafted_mu.json.gz", "MU"),
         ("edge_ms.json.gz", "MS"), ("edge_mu.json.gz", "MU")]


@pytest.mark.parametrize("name,_typ", PULSE)
def test_oracle_matches_reference_pulse(oracle, name, _typ):
    recs = load_golden(name)
    for typ in ("MS", "MU"):
        sel = [r for r in recs if r["type"] == typ]
        if not sel:
            continue
        batch = pack.pack_pulse([r["msg"] for r in sel], pack.KIND_BY_NAME[typ])
        got = oracle.run_pulse(batch)
        exp = [golden_expected(r) for r in sel]
        assert got == exp, diff_report(got, exp)


@pytest.mark.parametrize("name,kind,repaired", [("corpus_mc_strict.json.gz", pack.KIND_MC, False),
                                                ("corpus_mc_repaired.json.gz", pack.KIND_MC, True),
                                                ("corpus_mn.json.gz", pack.KIND_MN, True)])
def test_oracle_matches_reference_hex(oracle, protocols, name, kind, repaired):
    recs = load_golden(name)
    index = {pid: i for i, pid in enumerate(protocols)}
    batch = pack.pack_hex([r["msg"] for r in recs], kind, index)
    got = oracle.run_hex(batch, mc_repaired=repaired)
    exp = [golden_expected(r) for r in recs]
    assert got == exp, diff_report(got, exp)


def test_golden_has_raised_and_multi_hit_cases():
    """The fixtures cover what the reference tests do not: exceptions and repeated frames."""
    mu = load_golden("corpus_mu.json.gz")
    assert any(r["status"] == "IndexError" for r in mu)
    assert any(len(r["results"]) >= 4 for r in mu)
    crafted = load_golden("crafted_mu.json.gz")          # > 4 matches per survivor, > 64 per message, empty captures
    assert sum(1 for r in crafted if len(r["results"]) > 64) >= 5
    assert any(r["status"] == "IndexError" for r in crafted)
    # adversarial sets that reach ACCEPT paths: corpus frames pushed to tolerance edges, duplicate candidates, permuted slots
    for name in ("edge_ms.json.gz", "edge_mu.json.gz"):
        edge = load_golden(name)
        assert sum(1 for r in edge if r["results"]) > 0.2 * len(edge)
        assert sum(1 for r in edge if not r["results"]) > 0.1 * len(edge)
    mc = load_golden("corpus_mc_strict.json.gz")
    assert all(r["status"] in ("ok", "TypeError") and not r["results"] for r in mc)
    assert any(r["status"] == "TypeError" for r in mc)
