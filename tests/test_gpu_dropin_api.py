"""Drop-in API: the calls the reference's own tests make, against the GPU-backed class.

Mirrors tests/test_ms_demodulation.py, tests/test_mu_demodulation.py, tests/test_ms_parser.py (engine part),
tests/test_sd_protocols.py and tests/test_helpers.py (pure helpers) of the reference; expected values are the
reference outputs captured in SURVEY.md App. C / tests/golden.
"""
import pytest

from pysignalduino_b200 import SDProtocols

pytestmark = pytest.mark.gpu


def parse_line(line):
    d = {}
    for part in line.split(";"):
        if not part:
            continue
        if "=" in part:
            k, v = part.split("=", 1)
            d[k] = v
        else:
            d[part] = ""
    if "D" in d:
        d["data"] = d["D"]
    return d


@pytest.fixture(scope="module")
def protocols():
    return SDProtocols()


def test_ms_protocol_3_1(protocols):
    msg = {"P0": "330", "P1": "-14520", "P2": "-1254", "P3": "1155", "P4": "-330", "data": "01" + "02" * 23 + "34",
           "CP": "0", "SP": "0", "R": "0"}
    results = protocols.demodulate(msg, "MS")
    assert results == [{"protocol_id": "3.1", "payload": "i000001", "meta": {"bit_length": 24, "rssi": "0", "clock": 330.0}}]


def test_ms_cul_tcm_97001(protocols):
    msg = parse_line("MS;P1=502;P2=-9212;P3=-1939;P4=-3669;D=12131413141414131313131313141313131313131314141414141413131313141413131413;CP=1;SP=2;")
    results = protocols.demodulate(msg, "MS")
    assert [r["protocol_id"] for r in results] == ["0.4", "0.3", "0"]
    assert all(r["payload"] == "s5C080FC32000" and r["meta"] == {"bit_length": 40, "rssi": None, "clock": 502.0} for r in results)


def test_ms_invalid_inputs_return_empty_and_log(protocols):
    logged = []
    protocols.register_log_callback(lambda m, lvl: logged.append((m, lvl)))
    bad = parse_line("MS;P1=-8043;P2=505;P3=-1979;P4=-3960;D=21212323;CP=2;SP=1;R=1q;")
    assert protocols.demodulate(bad, "MS") == []
    assert protocols.demodulate({"data": "12a4", "CP": "1", "SP": "2"}, "MS") == []
    assert protocols.demodulate({"data": "1234", "CP": "", "SP": "2"}, "MS") == []
    assert protocols.demodulate({}, "XX") == []
    assert len(logged) == 4 and all(lvl == 3 for _, lvl in logged)
    protocols._log_callback = None


def test_mu_protocol_44(protocols):
    msg = parse_line("MU;P0=32001;P1=-1939;P2=1967;P3=3896;P4=-3895;D=01213424242124212121242121242121212124212424212121212121242421212421242121242124242421242421242424242124212124242424242421212424212424212121242121212;CP=2;R=39;")
    results = protocols.demodulate(msg, "MU")
    assert results == [{"protocol_id": "44", "payload": "W44#D12160652EDE9F9B10", "meta": {"bit_length": 72, "rssi": "39", "clock": 500.0}}]


def test_mu_protocol_46_repeats(protocols):
    msg = parse_line("MU;P0=-1943;P1=1966;P2=-327;P3=247;P5=-15810;D=01230121212301230121212121230121230351230121212301230121212121230121230351230121212301230121212121230121230351230121212301230121212121230121230351230121212301230121212121230121230351230;CP=1;")
    results = protocols.demodulate(msg, "MU")
    assert len(results) == 4 and all(r["protocol_id"] == "46" and r["payload"] == "P46#BAFB0" for r in results)
    assert results[0]["meta"] == {"bit_length": 20, "rssi": None, "clock": 290.0}


def test_mu_index_error_propagates(protocols):
    msg = parse_line("MU;P0=480;P1=-960;P2=-480;CP=0;D=0102010101010102020101020201010202010101020202010202020201020101010101010201020202010201010101010101020102010201020201010202010102010201020201")
    with pytest.raises(IndexError):
        protocols.demodulate(msg, "MU")
    statuses, results = protocols.demodulate_batch([msg], "MU")
    assert statuses == ["IndexError"] and results == [[]]


def test_mu_postdemod_valueerror_is_swallowed(protocols):
    msg = {"P0": "366", "P1": "-854", "P2": "854", "P3": "-366", "data": "23" * 40 + "01", "CP": "0"}
    out = protocols.demodulate(msg, "MU")
    assert {"protocol_id": "60", "payload": "K00000000008", "meta": {"bit_length": 44, "rssi": None, "clock": 122.0}} in out


def test_mc_as_shipped_raises_typeerror_and_repaired_decodes():
    msg = {"protocol_id": "43", "data": "A1B2C3D4E5F6A7", "clock": 640, "bit_length": 56}
    with pytest.raises(TypeError):
        SDProtocols(mc_repaired=False).demodulate(dict(msg), "MC")
    out = SDProtocols(mc_repaired=True).demodulate(dict(msg), "MC")
    assert out == [{"protocol_id": "43", "payload": "YsA1B2C3D4E5F6A7", "meta": {"protocol_id": "43", "rssi": None, "freq_afc": None}}]
    s = SDProtocols(mc_repaired=True)
    assert s.demodulate({"protocol_id": "10", "data": "AAAAAAAAAAAAAAAA3", "clock": 400, "bit_length": 68}, "MC")[0]["payload"] == "5555555555555555C"
    assert s.demodulate({"protocol_id": "119", "data": "9D4F3F7555A00", "clock": 500, "bit_length": 52}, "MC")[0]["payload"] == "J2C175F30008F"
    assert s.demodulate({"protocol_id": "119", "data": "9D4F3F7555A00", "clock": 520, "bit_length": 52}, "MC") == []
    assert s.demodulate({"protocol_id": "96", "data": "AAAAAAAA", "clock": 200, "bit_length": 49}, "MC")[0]["payload"] == "P96#AAAAAAAA"
    assert s.demodulate({"data": "AAAA", "clock": 1, "bit_length": 1}, "MC") == []
    with pytest.raises(TypeError):
        s.demodulate({"protocol_id": "57", "data": "AAAAAA", "clock": 330, "bit_length": 22}, "MC")   # helpers.mcraw int > str


def test_mn_converters(protocols):
    r = protocols.demodulate({"protocol_id": "115", "data": "3BF120B00C1618FF77FF0458152293FFF06B0000"}, "MN")
    assert r == [{"protocol_id": "115", "payload": "3BF120B00C1618FF77FF0458152293FFF06B0000", "meta": {}}]
    r = protocols.demodulate({"protocol_id": "100", "data": "9AA6362CC8AAAA000012F8F4"}, "MN")
    assert r == [{"protocol_id": "100", "payload": "OK 9 42 129 4 212 44", "meta": {"is_raw": False}}]
    assert protocols.demodulate({"protocol_id": "107", "data": "AABB"}, "MN") == []
    assert protocols.demodulate({"data": "AABB"}, "MN") == []


def test_property_api_and_table_mutation(protocols):
    assert len(protocols.get_keys("sync")) == 66 and len(protocols.get_keys("clockabs")) == 129
    assert protocols.protocol_exists("3.1") and not protocols.protocol_exists("9999")
    assert protocols.check_property("9", "clockabs") == 480 and protocols.get_property("9", "nope") is None
    assert protocols.check_property("9", "active") is True and protocols.get_property("9", "name")
    s = SDProtocols()
    msg = parse_line("MU;P0=-28704;P1=450;P2=-1064;P3=1422;CP=1;R=13;D=0121212121212121232121212121212121212121232323232321232123212321232323232323232323232323232323232323232323232323232321212121232101212121212121212321212121212121212121212323232323212321232123212323232323232323232323232323232323232323232323232323212121212321;")
    before = [r["protocol_id"] for r in s.demodulate(msg, "MU")]
    assert "9" in before
    s._protocols["9"]["active"] = False                     # reference honours `active` in MU (message_unsynced.py:48)
    after = [r["protocol_id"] for r in s.demodulate(msg, "MU")]
    assert "9" not in after and [p for p in before if p != "9"] == after


def test_pure_helpers(protocols):
    assert protocols.bin_str_2_hex_str("1111") == "F" and protocols.bin_str_2_hex_str("10011") == "13"
    assert protocols.bin_str_2_hex_str("") == "" and protocols.bin_str_2_hex_str("10F1") is None
    assert protocols.hex_to_bin_str("00FF") == "11111111" and protocols.hex_to_bin_str("0000") == "0000"
    assert protocols.hex_to_bin_str("zz") is None
    assert protocols.length_in_range("9", 60) == (1, "") and protocols.length_in_range("9", 1)[0] == 0
    assert protocols.length_in_range("nope", 1) == (0, "protocol does not exists")
    assert protocols.dec_2_bin_ppari(32) == "001000001"
    assert protocols.mc2dmc("1010") == "000"


def test_demod_host_payloads_equals_decode_then_format(sdp):
    """sdb_demod_host_payloads (decode + payload strings; MS / MU strings come from the device format kernel, stage by stage)
    returns exactly what sdb_demod_host followed by the host formatter sdb_format_hits returns — multi-stage (pipelined) and
    single-stage batches, all four kinds, with and without the bit arena, and reports the pool size it needs."""
    import numpy as np

    from corpus.corpus import Corpus
    from pysignalduino_b200 import pack

    corpus = Corpus(sdp.get_protocol_list())
    eng = sdp.engine()
    for kind, n in ((pack.KIND_MU, 300_000), (pack.KIND_MS, 270_000), (pack.KIND_MU, 1000), (pack.KIND_MC, 5000), (pack.KIND_MN, 5000)):
        b = corpus.pulse(kind, n) if kind <= 1 else corpus.hexmsgs(kind, n)
        ref = eng.demod_host(b, mc_repaired=True)
        pool_ref, off_ref = eng.format_hits(kind, ref.hits, ref.bits)
        pool_ref = np.frombuffer(pool_ref, dtype=np.uint8)
        nh = len(ref.hits)
        out = np.zeros(n, dtype=pack.MSGOUT_DTYPE)
        phits = np.zeros(nh + 8, dtype=pack.PAYHIT_DTYPE)
        ctr = np.zeros(1, dtype=pack.COUNTERS_DTYPE)
        pool = np.zeros(len(pool_ref) + nh + 16, dtype=np.uint8)
        # the 16-byte hit records and the bit arena are optional extras
        hits = np.zeros(nh + 8, dtype=pack.HIT_DTYPE) if kind == pack.KIND_MN else None
        bits = np.zeros(len(ref.bits) + 8, dtype=np.uint32) if kind == pack.KIND_MN else None
        args = (kind, np.ascontiguousarray(b.msgs), np.ascontiguousarray(b.digits), out, phits, ctr)
        kw = dict(mc_repaired=True, bits_cap=len(ref.bits) + 8, hits=hits, bits=bits)
        rc, used = eng.demod_host_payloads_into(*args, pool, **kw)
        assert rc == 0 and int(ctr["hits"][0]) == nh and used == len(pool_ref) + nh          # one NUL per hit
        if hits is not None:
            assert np.array_equal(hits["proto"][:nh], phits["proto"][:nh]) and np.array_equal(hits["nbits"][:nh], phits["nbits"][:nh])

        # hit order differs between runs (atomics), so compare per message: the strings of message m in hit order
        def strings(o, k, get):
            res = {}
            for m in np.nonzero(o["nhits"])[0][:: max(1, n // 3000)]:
                h0, c = int(o["hit_off"][m]), int(o["nhits"][m])
                res[int(m)] = [get(i) for i in range(h0, h0 + c)]
            return res

        def nul_string(i):
            a = int(phits["off"][i])
            e = a
            while pool[e]:
                e += 1
            return bytes(pool[a:e])

        got = strings(out, kind, nul_string)
        exp = strings(ref.out, kind, lambda i: bytes(pool_ref[int(off_ref[i]) : int(off_ref[i + 1])]))
        assert got == exp and len(got) > 100
        small = np.zeros(8, dtype=np.uint8)
        rc, used2 = eng.demod_host_payloads_into(*args, small, **kw)
        assert rc == -3 and used2 == used                       # SDB_E_OVERFLOW reports the size needed
        res2, pool2 = eng.demod_payloads(b, mc_repaired=True)
        assert len(res2.hits) == nh and len(pool2) == used and res2.hits.dtype == pack.PAYHIT_DTYPE


def test_handles_with_different_tables_coexist(sdp):
    """The resolve kernels keep table copies in dynamic shared memory whose size depends on the table; the opt-in attribute is per
    kernel function, not per handle: a handle created later with a much smaller table must not break an earlier one."""
    from corpus.corpus import Corpus
    from pysignalduino_b200 import pack

    b = Corpus(sdp.get_protocol_list()).pulse(pack.KIND_MU, 2000)
    before = sdp.demodulate_packed(b)
    small = SDProtocols(device=0)
    keep = ["9", "44", "46"]
    for pid in [p for p in list(small._protocols) if p not in keep]:
        del small._protocols[pid]
    msg = parse_line("MU;P0=-28704;P1=450;P2=-1064;P3=1422;CP=1;R=13;D=012121212121212123212121212121212121212123232323232123212321232123232323232323232323232323232323232323232323232323232121212123210121212121212121232121212121212121212121232323232321232123212321232323232323232323232323232323232323232323232323232321212121232101212121212121212321212121212121212121212323232323212321232123212323232323232323232323232323232323232323232323232323212121212321;")
    assert [r["protocol_id"] for r in small.demodulate(msg, "MU")] == ["9", "9", "9"]
    after = sdp.demodulate_packed(b)                      # the big table's handle still launches
    assert int(after.counters["hits"]) == int(before.counters["hits"]) > 0
