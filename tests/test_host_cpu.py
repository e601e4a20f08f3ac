"""CPU-only checks of the host logic: packer, table compiler, C ABI exports, scalar oracle pieces."""
import ctypes
import math
import random
import re

import numpy as np
import pytest

from pysignalduino_b200 import pack, table
from pysignalduino_b200.capi import EXPORTS, LIB_PATH


def test_library_exports_every_declared_symbol():
    """libsdb200.so loads without a GPU and exports every function include/sdb200.h declares."""
    from pysignalduino_b200 import build_ext

    build_ext.build()
    lib = ctypes.CDLL(str(LIB_PATH))
    header = (LIB_PATH.parent.parent / "include" / "sdb200.h").read_text()
    declared = set(re.findall(r"\b(sdb_[a-z_]+)\s*\(", header))
    assert declared == set(EXPORTS)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.sdb_abi_version() == 3


def test_create_fails_loudly_without_gpu():
    """No CPU fallback: without a CUDA device the engine cannot be created."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from pysignalduino_b200 import SDProtocols
    from pysignalduino_b200.capi import SdbError

    with pytest.raises(SdbError):
        SDProtocols().demodulate({"data": "0101", "P0": "300", "P1": "-300"}, "MU")


def test_record_layouts_match_header():
    assert pack.PULSE_DTYPE.itemsize == 48 and pack.HEX_DTYPE.itemsize == 16
    assert pack.MSGOUT_DTYPE.itemsize == 8 and pack.HIT_DTYPE.itemsize == 16 and pack.COUNTERS_DTYPE.itemsize == 16
    assert table.PULSEPROTO_DTYPE.itemsize == 248 and table.KEYTPL_DTYPE.itemsize == 48
    assert table.PREFILTER_DTYPE.itemsize == 28 and table.MMITEM_DTYPE.itemsize == 20 and table.HEXPROTO_DTYPE.itemsize == 36


def test_pack_pattern_order_and_duplicates():
    """Slot order = dict insertion order; a later duplicate id overwrites the value, keeps the position
    (message_synced.py:50-57); unparsable values are skipped."""
    m = {"P3": "100", "P01": "-200", "P1": "-250", "P5": "", "P7": "x", "data": "3113", "CP": "3", "SP": "1"}
    b = pack.pack_pulse([m], pack.KIND_MS)
    r = b.msgs[0]
    assert int(r["npat"]) == 2
    assert [int(r["pat"][0]), int(r["pat"][1])] == [100, -250]
    assert [(int(r["pat_ids"]) >> 0) & 15, (int(r["pat_ids"]) >> 4) & 15] == [3, 1]
    assert int(r["cp"]) == 0 and b.clock[0] == 100.0


def test_pack_ms_gates():
    ok = {"P0": "300", "data": "0101", "CP": "0", "SP": "0"}
    bad = [dict(ok, data=""), dict(ok, data="01a1"), dict(ok, CP="x"), dict(ok, SP=""), dict(ok, R="1q"), dict(ok, R="")]
    b = pack.pack_pulse([ok] + bad, pack.KIND_MS)
    assert [int(f) & 1 for f in b.msgs["flags"]] == [1, 0, 0, 0, 0, 0, 0]


def test_pack_domain_errors():
    """Unrepresentable messages are marked per message (status DomainError on the device); strict=True raises."""
    bad = [{"P0": "300.5", "data": "00"}, {"P12": "300", "data": "00"}, {"P0": "300", "data": "0" * (pack.MAX_DIGITS + 1)},
           {f"P{k}": "100" for k in range(9)} | {"data": "00"}]
    good = {"P0": "300", "P1": "-300", "data": "0101"}
    for m in bad:
        with pytest.raises(pack.DomainError):
            pack.pack_pulse([m], pack.KIND_MU, strict=True)
    b = pack.pack_pulse([good] + bad + [good], pack.KIND_MU)
    assert sorted(b.domain) == [1, 2, 3, 4]
    assert [int(f) for f in b.msgs["flags"]] == [pack.MSG_VALID] + [pack.MSG_DOMAIN] * 4 + [pack.MSG_VALID]
    assert int(b.msgs["dlen"][5]) == 4 and int(b.msgs["doff"][5]) == 1          # the good messages keep their streams
    ok_long = pack.pack_pulse([{"P0": "300", "data": "0" * pack.MAX_DIGITS}], pack.KIND_MU)
    assert not ok_long.domain and int(ok_long.msgs["dlen"][0]) == pack.MAX_DIGITS
    hexbad = [{"protocol_id": "10", "data": "abcd", "clock": 400, "bit_length": 16},
              {"protocol_id": "10", "data": "A" * (pack.MAX_HEX + 1), "clock": 400, "bit_length": 16},
              {"protocol_id": "10", "data": "XYZ", "clock": 400, "bit_length": 16}]
    for m in hexbad:
        with pytest.raises(pack.DomainError):
            pack.pack_hex([m], pack.KIND_MC, {"10": 0}, strict=True)
    hb = pack.pack_hex(hexbad + [{"protocol_id": "10", "data": "ABCD", "clock": 400, "bit_length": 16}], pack.KIND_MC, {"10": 0})
    assert sorted(hb.domain) == [0, 1, 2] and [int(f) for f in hb.msgs["flags"]] == [pack.MSG_DOMAIN] * 3 + [pack.MSG_VALID]


def test_pack_unpack_roundtrip(corpus):
    for kind in (pack.KIND_MS, pack.KIND_MU):
        b = corpus.pulse(kind, 300)
        again = pack.pack_pulse([pack.unpack_pulse(b, i) for i in range(b.n)], kind)
        assert (again.msgs == b.msgs).all() and np.array_equal(again.digits, b.digits)


def test_non_digit_characters_become_other_nibble():
    b = pack.pack_pulse([{"P0": "300", "data": "0a1²"}], pack.KIND_MU)
    by = b.digits[:2]
    assert [by[0] & 15, by[0] >> 4, by[1] & 15, by[1] >> 4] == [0, 0xE, 1, 0xE]


def test_tenths_intervals_follow_reference_float_semantics():
    """intervals come from the reference's float expressions, not decimal arithmetic."""
    assert table.tenths_interval(-4.0)[:2] == (-51, -29)
    assert table.tenths_interval(1.2)[:2] == (2, 21)
    assert table.tenths_interval(-118.0)[:2] == (-1392, -968)
    lo, hi, ranks = table.tenths_interval(1.0)
    assert (lo, hi) == (0, 20) and ranks[10] == 0 and len(ranks) == 21


def test_tolerance_values():
    """tests/test_pattern_utils.py of the reference pins these."""
    assert table.calculate_tolerance(1) == 1.0 and table.calculate_tolerance(3) == 1.0
    assert table.calculate_tolerance(4) == pytest.approx(1.2) and table.calculate_tolerance(-10) == pytest.approx(3.0)
    assert table.calculate_tolerance(20) == pytest.approx(3.6)


def test_modulematch_compiler():
    items, end = table.compile_modulematch("^P114#[13569BDE][13579BDF]F$")
    assert end and len(items) == 8
    rest, never = table.fold_preamble(items, end, "P114#")
    assert not never and len(rest) == 3
    rest, never = table.fold_preamble(*table.compile_modulematch("^P15#.*"), "P15#")
    assert not never
    rest, never = table.fold_preamble(*table.compile_modulematch("^W64*"), "W64#")
    assert not never and rest == []
    rest, never = table.fold_preamble(*table.compile_modulematch("^P15#"), "P99#")
    assert never
    with pytest.raises(NotImplementedError):
        table.compile_modulematch("P15#")


def test_table_compiles_and_recompiles(protocols):
    ct = table.compile_table(protocols)
    assert ct.info["n_ms"] == 47 and ct.info["n_mu"] == 129 and ct.info["n_clk"] == 55
    assert len(ct.blob) == ct.info["bytes"] and len(ct.ids) == 160
    import copy

    p2 = copy.deepcopy(protocols)
    p2["9"]["active"] = False
    assert table.compile_table(p2).info["n_mu"] == 128


def test_oracle_round1_matches_cpython():
    """The oracle's fast round(x, 1) and the independent printf form equal CPython's round on ties and near-ties."""
    from oracle.oracle import lib

    L = lib()
    rng = random.Random(5)
    clocks = [244.0, 470.0, 406.0, 189.0, 122.0, -1.0, 100.0, 635.0, 333.0, 7.0]
    for _ in range(20000):
        p = rng.randint(-99999, 99999)
        c = rng.choice(clocks)
        x = p / c
        want = round(x, 1)
        assert L.ora_round1(x) == want and L.ora_round1_printf(x) == want
    for x in (0.25, 0.75, 0.35, 0.05, 0.15, -0.25, 2.5, -0.04, 1e-9):
        assert L.ora_round1(x) == round(x, 1)
        assert math.copysign(1, L.ora_round1(x)) == math.copysign(1, round(x, 1))
