"""The rest of the reference's API surface: loader, pattern_utils, by-name MC / MN internals, checksum helpers.

Reference-generated vectors (tests/golden/make_golden_api.py -> api_units.json.gz) plus ports of the reference's own
tests: tests/test_loader.py:2-40, tests/test_pattern_utils.py:5-95, tests/test_manchester_protocols.py:51-100,
tests/test_rsl_handler.py.  Host-only helpers run without a GPU; everything that decodes is marked gpu.
"""
import pytest

from pysignalduino_b200 import SDProtocols, pack
from tests.common import load_golden


@pytest.fixture(scope="module")
def gold():
    import gzip
    import json
    from tests.common import GOLDEN

    with gzip.open(GOLDEN / "api_units.json.gz", "rt", encoding="utf-8") as f:
        return json.load(f)


def _same(result, rec):
    """Compare a call outcome with a golden {"ok": value} / {"raises": name} record (tuples became lists in JSON)."""
    def norm(x):
        if isinstance(x, (tuple, list)):
            return [norm(v) for v in x]
        if isinstance(x, dict):
            return {k: norm(v) for k, v in x.items()}
        return x
    return norm(result) == norm(rec)


def _call(fn, *a, **k):
    try:
        return {"ok": fn(*a, **k)}
    except pack.DomainError:
        raise
    except Exception as e:  # noqa: BLE001
        return {"raises": type(e).__name__}


# ------------------------------------------------------------------------------------------ host-only surface
def test_checksum_helpers_match_reference(gold):
    s = SDProtocols()
    for r in gold["lfsr_digest16"]:
        assert _same(_call(s.lfsr_digest16, *r["args"]), r["result"]), r
    for r in gold["calc_crc16"]:
        assert _same(_call(s._calc_crc16, *r["args"]), r["result"]), r
    for r in gold["calc_crc8_la_crosse"]:
        assert _same(_call(s._calc_crc8_la_crosse, *r["args"]), r["result"]), r
    for r in gold["mc_hex_bits"]:
        assert _same(_call(s._convert_mc_hex_to_bits, "gold", r["raw_hex"], r["invert"], len(r["raw_hex"])), r["result"]), r


def test_rsl_placeholders():
    """tests/test_rsl_handler.py: the upstream placeholders echo their input."""
    s = SDProtocols()
    assert s.decode_rsl("1010101010") == {"decoded": "1010101010", "status": 1}
    assert s.encode_rsl({"a": 1}) == {"encoded": "{'a': 1}", "status": 1}


def test_loader_resolves_methods():
    """tests/test_loader.py:12-31 (the parts that need no device)."""
    from pysignalduino_b200.loader import protocols, resolve_method, run_method

    found = [p["method"] for p in protocols.values() if p.get("method") in ("manchester.mcBit2Grothe", "manchester.mcBit2SomfyRTS")]
    assert found
    for path in ("manchester.mcBit2Grothe", "manchester.mcBit2SomfyRTS", "rsl_handler.decode_rsl", "rsl_handler.encode_rsl",
                 "helpers.ConvBresser_6in1", "helpers.mcraw", "postdemodulation.postDemo_EM"):
        assert callable(resolve_method(path))
    assert resolve_method("rsl_handler.decode_rsl")("1010101010") is not None
    with pytest.raises(ValueError):
        resolve_method("mcBit2Grothe")
    with pytest.raises(AttributeError):
        resolve_method("manchester.no_such_method")
    with pytest.raises(ValueError):
        run_method("9", "x")                       # protocol 9 has no method
    with pytest.raises(ValueError):
        run_method("nope", "x")
    assert "active" not in protocols["9"]          # the loader's table is the raw JSON (no defaults applied)


def test_pattern_utils_host_functions():
    """tests/test_pattern_utils.py:7-14"""
    from pysignalduino_b200.pattern_utils import calculate_tolerance, cartesian_product, is_in_tolerance

    assert calculate_tolerance(1) == 1.0 and calculate_tolerance(3) == 1.0
    assert calculate_tolerance(4) == pytest.approx(1.2) and calculate_tolerance(10) == pytest.approx(3.0)
    assert calculate_tolerance(20) == pytest.approx(3.6) and calculate_tolerance(-10) == pytest.approx(3.0)
    assert is_in_tolerance(1.0, 1.5, 0.5) and not is_in_tolerance(1.0, 1.6, 0.5)
    assert cartesian_product([]) == [[]] and cartesian_product([[1, 2], [3]]) == [[1, 3], [2, 3]]


def test_protocol_dict_mutation_is_tracked():
    """Edits of the table — outer dict, rows, lists inside rows — bump one version counter (tracked.py); the compiled table
    follows without a per-call hash of the dict."""
    import copy
    import json

    s = SDProtocols()
    v0 = s._ver.n
    t0 = s.compiled_table()
    assert s.compiled_table() is t0
    s._protocols["9"]["active"] = False
    assert s._ver.n > v0 and s.compiled_table() is not t0 and "9" not in s.compiled_table().mu_ids
    v1 = s._ver.n
    s._protocols["9"]["one"][0] = 2                       # a list inside a row
    assert s._ver.n > v1
    v2 = s._ver.n
    s._protocols["9"] = {"length_min": 50, "name": "TestLength"}      # tests/test_manchester_protocols.py:54
    s._protocols["9"]["length_max"] = 60
    assert s._ver.n >= v2 + 2 and s.length_in_range("9", 61) == (0, "message is too long")
    s.get_protocol_list().pop("9")
    assert not s.protocol_exists("9")
    assert isinstance(s._protocols["10"], dict) and isinstance(s._protocols["10"]["clockrange"], list)
    json.dumps(s._protocols)
    c = copy.deepcopy(s._protocols)
    c["10"]["name"] = "x"
    assert s._protocols["10"]["name"] != "x"
    s._protocols = {"1": {"name": "only"}}                # wholesale replacement
    assert s.get_keys() == ["1"]


def test_table_compiler_degrades_per_protocol(protocols):
    """A user-edited table with shapes the device layout cannot hold compiles anyway (the reference skips what it cannot use,
    message_synced.py:206, message_unsynced.py:234): the row is left out and reported, a free-form modulematch is handed to the
    host formatter; strict=True keeps the hard failure."""
    import copy

    from pysignalduino_b200 import table

    t = copy.deepcopy(protocols)
    t["9"]["one"] = [1, -2, 3, -4, 5]                    # 5 distinct values + symbol width 5
    t["44"]["modulematch"] = "^W44#(AA|BB)+.*$"         # alternation: not a fixed-offset class program
    t["3"]["start"] = [1.0] * 20                         # template longer than 14 pulses
    t["13"]["length_min"] = "abc"                        # not a number
    c = table.compile_table(t)
    assert {"9 (MU)", "3 (MU)", "13 (MS)"} <= set(c.unsupported) and "44 (MU)" not in c.unsupported
    assert "9" not in c.mu_ids and "44" in c.mu_ids and len(c.mu_ids) >= 125
    rows = [r for r in __import__("numpy").frombuffer(c.blob, dtype=table.PULSEPROTO_DTYPE, count=c.info["n_mu"],
                                                      offset=int(__import__("numpy").frombuffer(c.blob, dtype=table.HEADER_DTYPE, count=1)[0]["off_mu"]))]
    r44 = next(r for r in rows if c.ids[int(r["proto"])] == "44")
    assert int(r44["flags"]) & table.PF_MM_HOST
    with pytest.raises((NotImplementedError, ValueError)):
        table.compile_table(t, strict=True)
    assert not table.compile_table(protocols).unsupported          # the shipped table compiles completely


# ------------------------------------------------------------------------------------------ device-backed surface
@pytest.mark.gpu
def test_loader_runs_methods():
    """tests/test_loader.py:33-53: run every selected method once through the loader."""
    from pysignalduino_b200.loader import protocols, resolve_method, run_method

    for func in ("mcBit2Grothe", "mcBit2SomfyRTS"):
        pid = next(p for p, d in protocols.items() if d.get("method") == f"manchester.{func}")
        assert resolve_method(f"manchester.{func}")(name="test", bit_data="1010101010", protocol_id=pid, mcbitnum=10) is not None
    pid = next(p for p, d in protocols.items() if d.get("method") == "manchester.mcBit2Grothe")
    assert run_method(pid, "test", "10101010101010101010101010101010", pid, 32) == (1, "AAAAAAAA")


@pytest.mark.gpu
def test_pattern_exists_reference_cases():
    """tests/test_pattern_utils.py:15-95 on the device resolver."""
    from pysignalduino_b200.pattern_utils import pattern_exists

    assert pattern_exists([1, -1], {"0": 1.0, "1": -1.0}, "0101") == "01"
    assert pattern_exists([10, -5], {"0": 11.0, "1": -4.0}, "01") == "01"
    assert pattern_exists([1], {"0": 20.0}, "0") == -1
    assert pattern_exists([1], {"0": 1.0}, "222") == -1
    assert pattern_exists([1, 2], {"0": 1.5}, "00") == -1                      # one id cannot stand for two values
    assert pattern_exists([1, 1], {"0": 1.0}, "00") == "00"
    assert pattern_exists([1], {"0": 1.0, "1": 1.1}, "1") == "1"              # second candidate by gap order
    with pytest.raises(pack.DomainError):
        pattern_exists([1], {"0": 1.04}, "0")
    with pytest.raises(pack.DomainError):
        pattern_exists([1], {"10": 1.0}, "0")


@pytest.mark.gpu
def test_pattern_exists_matches_reference(gold):
    from pysignalduino_b200.pattern_utils import pattern_exists

    bad, outside, found = [], 0, 0
    for i, r in enumerate(gold["pattern_exists"]):
        in_domain = len(set(float(v) for v in r["search"])) <= 4 and len(r["search"]) <= 14
        try:
            got = pattern_exists(r["search"], {k: v for k, v in r["patterns"]}, r["data"])
        except pack.DomainError:
            assert not in_domain, r                     # only what the module documents as outside the packed domain
            outside += 1
            continue
        found += got != -1
        if got != r["result"]:
            bad.append((i, r["search"], r["patterns"], r["data"][:40], got, r["result"]))
    assert not bad, bad[:5]
    assert found > 200 and outside < len(gold["pattern_exists"]) // 5


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["strict", "repaired"])
def test_demodulate_mc_data_matches_reference(gold, mode):
    """_demodulate_mc_data (manchester.py:49-144) incl. the edited-row cases of tests/test_manchester_protocols.py:51-100."""
    s = SDProtocols(device=0, mc_repaired=mode == "repaired")
    bad = []
    for i, r in enumerate(gold["mc_data"]):
        saved = None
        if r["table"] is not None:
            saved = s._protocols["119"]
            s._protocols["119"] = dict(r["table"])
        got = _call(s._demodulate_mc_data, **r["args"])
        if saved is not None:
            s._protocols["119"] = saved
        if not _same(got, r[mode]):
            bad.append((i, r["args"], r["table"], got, r[mode]))
    assert not bad, bad[:5]


@pytest.mark.gpu
def test_mc_demodulate_length_check_port():
    """tests/test_manchester_protocols.py:51-100 verbatim in behaviour."""
    proto = SDProtocols(device=0)
    pid = "119"
    proto._protocols[pid] = {"length_min": 50, "name": "TestLength"}
    kw = dict(name="TestLen", protocol_id=pid, clock=500, raw_hex="AABBCCDD1122", mcbitnum=48, messagetype="MC", version=None)
    result = proto._demodulate_mc_data(**kw)
    assert result[0] == -1 and result[1] == "message is too short"
    proto._protocols[pid]["length_min"] = 10
    proto._protocols[pid]["length_max"] = 40
    proto._protocols[pid]["method"] = "manchester.mcRaw"
    result = proto._demodulate_mc_data(**kw)
    assert result[0] == -1 and result[1] == "message is too long"


@pytest.mark.gpu
def test_demodulate_mn_data_and_mcraw_match_reference(gold, sdp):
    for r in gold["mn_data"]:
        assert _same(_call(sdp._demodulate_mn_data, "gold", r["protocol_id"], r["msg"]), r["result"]), r
    for r in gold["mcraw"]:
        assert _same(_call(sdp.mcraw, *r["args"]), r["result"]), r


@pytest.mark.gpu
def test_host_evaluated_modulematch(sdp):
    """A modulematch the device program cannot express is applied by the host formatter (PF_MM_HOST / SDB_HIT_MM_HOST): same
    hits as an equivalent expressible regex, and a never-matching one removes them."""
    r = next(x for x in load_golden("reference_vectors.json.gz") if any(h[0] == "44" for h in x["results"]))
    s = SDProtocols(device=0)
    base = s.demodulate(r["msg"], "MU")
    assert any(o["protocol_id"] == "44" for o in base)
    s._protocols["44"]["modulematch"] = "^W44#(D1|XX)[0-9A-F]+$"
    assert "44 (MU)" not in s.compiled_table().unsupported
    assert s.demodulate(r["msg"], "MU") == base
    s._protocols["44"]["modulematch"] = "^W44#(XX|YY)"
    assert [o for o in s.demodulate(r["msg"], "MU")] == [o for o in base if o["protocol_id"] != "44"]


@pytest.mark.gpu
def test_scalar_call_is_not_slower_than_a_table_hash(sdp):
    """The per-call cost of noticing table edits is one integer compare: 200 scalar MS calls stay far below the 0.9 ms per
    call that hashing the table cost (the reference's whole MS decode takes ~0.5 ms)."""
    import time

    r = load_golden("reference_vectors.json.gz")[0]
    sdp.demodulate(r["msg"], r["type"])
    t0 = time.perf_counter()
    for _ in range(200):
        sdp.demodulate(r["msg"], r["type"])
    per_call = (time.perf_counter() - t0) / 200
    assert per_call < 0.9e-3, per_call
