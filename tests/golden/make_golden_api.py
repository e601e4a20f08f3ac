#!/usr/bin/env python3
"""Golden vectors for the remaining API surface (build container only): api_units.json.gz.

Direct calls of the REAL reference's by-name methods and module functions that the decode path, the loader and the
reference's own tests reach:
  pattern_utils.pattern_exists (pattern_utils.py:34-136)          random in-domain calls + the six cases of
                                                                  tests/test_pattern_utils.py:15-95
  SDProtocols._demodulate_mc_data (manchester.py:49-144)          as shipped and with the two documented repairs
  SDProtocols._demodulate_mn_data (manchester.py:147-204)
  SDProtocols._convert_mc_hex_to_bits (manchester.py:18-47)
  SDProtocols.lfsr_digest16 / _calc_crc16 / _calc_crc8_la_crosse  (helpers.py:190-221, :281-380)
  SDProtocols.mcraw with non-binary bit data (helpers.py:90-122)
"""
import gzip
import json
import random
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))
from corpus.corpus import Corpus, batch_to_dicts  # noqa: E402
from oracle import ref_import  # noqa: E402
from pysignalduino_b200 import pack  # noqa: E402
from pysignalduino_b200.protocol_data import load_protocol_table  # noqa: E402


def call(fn, *a, **k):
    try:
        return {"ok": fn(*a, **k)}
    except Exception as e:  # noqa: BLE001 - the exception type IS the result
        return {"raises": type(e).__name__}


def pattern_cases(rng, protocols):
    ref_import.reference_class()
    from sd_protocols.pattern_utils import pattern_exists  # type: ignore

    cases = [
        ([1, -1], [["0", 1.0], ["1", -1.0]], "0101"), ([10, -5], [["0", 11.0], ["1", -4.0]], "01"), ([1], [["0", 20.0]], "0"),
        ([1], [["0", 1.0]], "222"), ([1, 2], [["0", 1.5]], "00"), ([1, 1], [["0", 1.0]], "00"), ([1], [["0", 1.0], ["1", 1.1]], "1"),
        ([], [["0", 1.0]], "0"), ([1, -1], [], "01"), ([1, -2], [["0", 1.0], ["1", -2.0]], ""),
    ]
    templates = []
    for pr in protocols.values():
        for key in ("sync", "start", "one", "zero", "float"):
            v = pr.get(key)
            if isinstance(v, list) and v and len(v) <= 14:
                templates.append([float(x) for x in v])
    for _ in range(900):
        tpl = list(rng.choice(templates))
        if rng.random() < 0.2:
            tpl = [rng.choice([1, -1, 2, -2, 3.5, -4, 8, -9.5, 17, -31, 0.5]) for _ in range(rng.randrange(1, 6))]
        npat = rng.randrange(1, 9)
        ids = rng.sample("0123456789", npat)
        uniq = sorted(set(tpl))
        pats = []
        for pid in ids:
            if rng.random() < 0.75:
                base = rng.choice(uniq)
                spread = max(1.0, abs(base) * 0.35)
                val = round(base + rng.uniform(-spread, spread), 1)
            else:
                val = round(rng.uniform(-40, 40), 1)
            pats.append([pid, val])
        # data: mostly built from plausible id strings so that "found" happens, with noise
        L = rng.choice([0, 1, 2, 5, 30, 30, 60, 60, 200, 200, 200, 1024, 1500, 4096]) if rng.random() < 0.15 else rng.choice([2, 5, 30, 60, 120])
        alphabet = "".join(ids) + ("" if rng.random() < 0.7 else "0123456789x")
        data = "".join(rng.choice(alphabet) for _ in range(L))
        cases.append((tpl, pats, data))
    out = []
    for tpl, pats, data in cases:
        res = pattern_exists(list(tpl), {k: v for k, v in pats}, data)
        out.append({"search": tpl, "patterns": pats, "data": data, "result": res})
    return out


def mc_cases(rng, protocols):
    strict = ref_import.reference_class()()
    repaired = ref_import.repaired_class()()
    corp = Corpus(protocols)
    out = []
    msgs = batch_to_dicts(corp.hexmsgs(pack.KIND_MC, 600))
    for m in msgs:
        mt = "Mc" if rng.random() < 0.1 else "MC"
        ver = rng.choice([None, None, "V 3.2.0-dev", "V 3.5.0"])
        args = dict(name="gold", protocol_id=m["protocol_id"], clock=m["clock"], raw_hex=m["data"], mcbitnum=m["bit_length"],
                    messagetype=mt, version=ver)
        out.append({"args": args, "table": None, "strict": call(strict._demodulate_mc_data, **args),
                    "repaired": call(repaired._demodulate_mc_data, **args)})
    # edited rows (tests/test_manchester_protocols.py:51-100 and variations): no clockrange, explicit limits / methods
    edits = [
        {"length_min": 50, "name": "TestLength"},
        {"length_min": 10, "length_max": 40, "method": "manchester.mcRaw", "name": "TestLength"},
        {"length_min": 10, "length_max": 60, "method": "manchester.mcRaw", "name": "T", "preamble": "X#"},
        {"length_min": 10, "length_max": 60, "method": "manchester.mcRaw", "polarity": "invert"},
        {"length_min": 10, "method": "manchester.nothing_there"},
        {"length_min": 10, "length_max": 60},
        {"length_min": 10, "length_max": 60, "method": "manchester.mcBit2Grothe", "preamble": "P96#"},
        {"length_min": 10, "length_max": 60, "method": "manchester.mcBit2Hideki", "preamble": "P12#", "clockrange": [300, 600]},
    ]
    for ed in edits:
        for raw_hex, mcb, clock in (("AABBCCDD1122", 48, 500), ("00AB", 16, 500), ("A8C233B53A3E0A0783", 71, 450), ("AAAAAAAA", 32, 420),
                                    ("0000", 16, 500), ("75F2A8C233B5", 48, 700)):
            args = dict(name="gold", protocol_id="119", clock=clock, raw_hex=raw_hex, mcbitnum=mcb, messagetype="MC", version=None)
            row = {}
            for cls, key in ((strict, "strict"), (repaired, "repaired")):
                saved = cls._protocols["119"]
                cls._protocols["119"] = dict(ed)
                row[key] = call(cls._demodulate_mc_data, **args)
                cls._protocols["119"] = saved
            out.append({"args": args, "table": ed, **row})
    # unknown / non-str ids take check_property's defaults
    for pid, mcb in (("nope", 48), (119, 48), ("nope", 10000)):
        args = dict(name="gold", protocol_id=pid, clock=500, raw_hex="AABB", mcbitnum=mcb, messagetype="MC", version=None)
        out.append({"args": args, "table": None, "strict": call(strict._demodulate_mc_data, **args),
                    "repaired": call(repaired._demodulate_mc_data, **args)})
    return out


def main():
    protocols = load_protocol_table()
    rng = random.Random(0xA91)
    ref = ref_import.reference_class()()
    corp = Corpus(protocols)
    rec = {"pattern_exists": pattern_cases(rng, protocols), "mc_data": mc_cases(rng, protocols)}
    mn = []
    for m in batch_to_dicts(corp.hexmsgs(pack.KIND_MN, 400)):
        mn.append({"protocol_id": m["protocol_id"], "msg": m, "result": call(ref._demodulate_mn_data, "gold", m["protocol_id"], m)})
    for pid in ("10", "57", "107", "nope"):      # an MC decoder, helpers.mcraw, no method, unknown id
        m = {"protocol_id": pid, "data": "9AA6362CC8AAAA000012F8F4"}
        mn.append({"protocol_id": pid, "msg": m, "result": call(ref._demodulate_mn_data, "gold", pid, m)})
    rec["mn_data"] = mn
    hexes = ["", "0", "00FF", "AB12", "ab12", "F" * 40, "0" * 9 + "1", "XYZ"] + ["".join(rng.choice("0123456789ABCDEF") for _ in range(rng.randrange(1, 60))) for _ in range(60)]
    rec["mc_hex_bits"] = [{"raw_hex": h, "invert": inv, "result": call(ref._convert_mc_hex_to_bits, "gold", h, inv, len(h))}
                          for h in hexes for inv in (False, True)]
    lf = []
    for _ in range(200):
        n = rng.randrange(0, 24)
        h = "".join(rng.choice("0123456789ABCDEFabcdef") for _ in range(rng.randrange(0, 60)))
        if rng.random() < 0.05:
            h = h[: len(h) // 2] + "G" + h[len(h) // 2:]
        gen, key = rng.choice([(0x8810, 0xABF9), (0x8810, 0xBA95), (rng.randrange(65536), rng.randrange(65536))])
        lf.append({"args": [n, gen, key, h], "result": call(ref.lfsr_digest16, n, gen, key, h)})
    rec["lfsr_digest16"] = lf
    crc = []
    for _ in range(200):
        h = "".join(rng.choice("0123456789ABCDEF") for _ in range(2 * rng.randrange(0, 24)))
        if rng.random() < 0.05:
            h += "Z"
        a = [h, rng.choice([0x1021, 0x8005, rng.randrange(65536)]), rng.choice([0, 0xFFFF, rng.randrange(65536)]), rng.random() < 0.5,
             rng.random() < 0.5, rng.choice([0, 0xFFFF, rng.randrange(65536)])]
        crc.append({"args": a, "result": call(ref._calc_crc16, *a)})
    rec["calc_crc16"] = crc
    rec["calc_crc8_la_crosse"] = [{"args": [h], "result": call(ref._calc_crc8_la_crosse, h)}
                                  for h in ["", "00", "9AA6362C", "FFFFFFFF", "12345678", "ABC", "zz"] +
                                  ["".join(rng.choice("0123456789ABCDEF") for _ in range(8)) for _ in range(40)]]
    raw = []
    for pid in ("57", "11", "18", "nope", "96"):
        for bits in ("1010", "10F1", "12", "1" * 30, "2" * 30, ""):
            for mcb in (None, 4, 30, 100):
                raw.append({"args": ["gold", bits, pid, mcb], "result": call(ref.mcraw, "gold", bits, pid, mcb)})
    rec["mcraw"] = raw
    path = HERE / "api_units.json.gz"
    with gzip.GzipFile(path, "wb", mtime=0) as gz:
        gz.write(json.dumps(rec, separators=(",", ":")).encode("utf-8"))
    for k, v in rec.items():
        print(k, len(v))
    pe = rec["pattern_exists"]
    print("pattern_exists found:", sum(1 for r in pe if r["result"] != -1), "of", len(pe))
    print("mc_data repaired ok:", sum(1 for r in rec["mc_data"] if isinstance(r["repaired"].get("ok"), tuple) and r["repaired"]["ok"][0] == 1),
          "strict raises:", sum(1 for r in rec["mc_data"] if "raises" in r["strict"]))
    print(f"api_units.json.gz: {path.stat().st_size} bytes")


if __name__ == "__main__":
    main()
