"""GPU: the CUDA path against reference-generated golden fixtures (dict in -> dict out, full drop-in path)."""
import pytest

from pysignalduino_b200 import pack
from tests.common import diff_report, golden_expected, load_golden

pytestmark = pytest.mark.gpu

FILES = ["reference_vectors.json.gz", "corpus_ms.json.gz", "corpus_mu.json.gz", "fuzz_ms.json.gz", "fuzz_mu.json.gz",
         "corpus_mn.json.gz", "crafted_mu.json.gz", "edge_ms.json.gz", "edge_mu.json.gz"]


def canon(statuses, results):
    return [(st, [(str(r["protocol_id"]), r["payload"], int(r["meta"].get("bit_length", -1))) for r in lst])
            for st, lst in zip(statuses, results)]


@pytest.mark.parametrize("name", FILES)
def test_gpu_matches_reference(sdp, name):
    recs = load_golden(name)
    for typ in ("MS", "MU", "MN"):
        sel = [r for r in recs if r["type"] == typ]
        if not sel:
            continue
        got = canon(*sdp.demodulate_batch([r["msg"] for r in sel], typ))
        exp = [golden_expected(r) for r in sel]
        assert got == exp, diff_report(got, exp)


@pytest.mark.parametrize("name,repaired", [("corpus_mc_strict.json.gz", False), ("corpus_mc_repaired.json.gz", True)])
def test_gpu_matches_reference_mc(name, repaired):
    from pysignalduino_b200 import SDProtocols

    s = SDProtocols(device=0, mc_repaired=repaired)
    recs = load_golden(name)
    got = canon(*s.demodulate_batch([r["msg"] for r in recs], "MC"))
    exp = [golden_expected(r) for r in recs]
    assert got == exp, diff_report(got, exp)


def test_scalar_demodulate_reference_vectors(sdp):
    """SDProtocols.demodulate(msg, type): same lists, same meta, same exceptions as the reference."""
    for r in load_golden("reference_vectors.json.gz"):
        if r["status"] != "ok":
            with pytest.raises({"IndexError": IndexError, "TypeError": TypeError, "ValueError": ValueError}[r["status"]]):
                sdp.demodulate(r["msg"], r["type"])
            continue
        out = sdp.demodulate(r["msg"], r["type"])
        assert [(o["protocol_id"], o["payload"], o["meta"]["bit_length"]) for o in out] == [tuple(x) for x in r["results"]]
        for o in out:
            assert o["meta"]["rssi"] == r["msg"].get("R")
            assert isinstance(o["meta"]["clock"], float)
