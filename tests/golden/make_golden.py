#!/usr/bin/env python3
"""Generate the committed golden fixtures by running the REAL reference (build container only).

    python tests/golden/make_golden.py

Writes gzip-JSON files under tests/golden/; each record is
    {"type": "MS"|"MU"|"MC"|"MN", "msg": {...parser dict...}, "status": "ok"|"IndexError"|..., "results": [[id, payload, bit_length], ...]}
Inputs: the reference's own test vectors (tests/test_ms_demodulation.py:31-41, tests/test_ms_parser.py:21-51,
tests/test_mu_demodulation.py:27-80, tests/test_decompress_payload.py:16,23), the first rows of the
synthetic corpora (seeds 0x5D01 / 0x5D02) and adversarial fuzz (oracle/validate_vs_reference.py).
"""
from __future__ import annotations

import gzip
import json
import random
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))

from corpus.corpus import Corpus, batch_to_dicts  # noqa: E402
from oracle import ref_import  # noqa: E402
from oracle.validate_vs_reference import fuzz_pulse  # noqa: E402
from pysignalduino_b200 import pack  # noqa: E402
from pysignalduino_b200.protocol_data import load_protocol_table  # noqa: E402


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


REFERENCE_VECTORS = [
    ("MS", {"P0": "330", "P1": "-14520", "P2": "-1254", "P3": "1155", "P4": "-330", "data": "01" + "02" * 23 + "34", "CP": "0", "SP": "0", "R": "0"}),
    ("MS", parse_line("MS;P1=502;P2=-9212;P3=-1939;P4=-3669;D=12131413141414131313131313141313131313131314141414141413131313141413131413;CP=1;SP=2;")),
    ("MS", parse_line("MS;P2=476;P3=-3894;P4=-977;P5=-1966;D=23242525242524252524242524242424242524252524252525252525252424252524242524;CP=2;SP=3;R=240;O;m0;")),
    ("MS", parse_line("MS;P1=-8043;P2=505;P3=-1979;P4=-3960;D=2121232323242424232423242323232323242324232424232324242323232323232323232323232323242423;CP=2;SP=1;R=12;")),
    ("MS", parse_line("MS;P1=-8043;P2=505;P3=-1979;P4=-3960;D=2121232323242424232423242323232323242324232424232324242323232323232323232323232323242423;CP=2;SP=1;R=1q;")),
    ("MS", parse_line("MS;=0;L=L=-1020;L=H=935;S=L=-525;S=H=444;D=354133323044313642333731303246303541423044364430;C==487;L==89;R==24;")),
    ("MS", parse_line("MS;P1=;L=L=-1015;L=H=944;S=L=-512;S=H=456;D=353531313436304235313330433137433244353036423130;C==487;L==89;R==45;")),
    ("MU", parse_line("MU;P0=32001;P1=-1939;P2=1967;P3=3896;P4=-3895;D=01213424242124212121242121242121212124212424212121212121242421212421242121242124242421242421242424242124212124242424242421212424212424212121242121212;CP=2;R=39;")),
    ("MU", parse_line("MU;P0=-1943;P1=1966;P2=-327;P3=247;P5=-15810;D=01230121212301230121212121230121230351230121212301230121212121230121230351230121212301230121212121230121230351230121212301230121212121230121230351230121212301230121212121230121230351230;CP=1;")),
    ("MU", parse_line("MU;P0=-21520;P1=235;P2=-855;P3=846;P4=620;P5=-236;P7=-614;D=012323232454545454545451717451717171745171717171717171717174517171745174517174517174545;CP=1;R=217;")),
    ("MU", parse_line("MU;P0=7944;P1=-724;P2=742;P3=241;P4=-495;P5=483;P6=-248;D=01212121343434345656343434563434345634565656343434565634343434343434345634345634345634343434343434343434345634565634345656345634343456563421212121343434345656343434563434345634565656343434565634343434343434563434563434563434343434343434343434345634565634;CP=3;R=47;")),
    ("MU", parse_line("MU;P0=-28704;P1=450;P2=-1064;P3=1422;CP=1;R=13;D=012121212121212123212121212121212121212123232323232123212321232123232323232323232323232323232323232323232323232323232121212123210121212121212121232121212121212121212121232323232321232123212321232323232323232323232323232323232323232323232323232321212121232101212121212121212321212121212121212121212323232323212321232123212323232323232323232323232323232323232323232323232323212121212321;")),
    ("MU", parse_line("MU;P0=480;P1=-960;P2=-480;CP=0;D=0102010101010102020101020201010202010101020202010202020201020101010101010201020202010201010101010101020102010201020201010202010102010201020201")),
    ("MU", parse_line("MU;P0=-2272;P1=228;P2=-356;P3=635;P4=-562;P5=433;D=012345234345252343452523434345252345234343434523434345252343452525252525234523452343452345252525;CP=5;R=4;P3=;L=L=-2864;L=H=2980;S=L=-1444;S=H=1509;D=354146333737463037;C==1466;L==32;R==9;")),
    ("MU", parse_line("MU;P0=-370;P1=632;P2=112;P3=-555;P4=428;P5=-780;P6=180;P7=-200;CP=4;R=77;")),
    ("MU", {"P0": "366", "P1": "-854", "P2": "854", "P3": "-366", "data": "23" * 40 + "01", "CP": "0"}),
]


def run(ref, items):
    out = []
    for typ, msg in items:
        st, res = ref_import.ref_demodulate(ref, msg, typ)
        out.append({"type": typ, "msg": msg, "status": st, "results": [list(r) for r in res]})
    return out


def dump(name, recs):
    path = HERE / name
    with gzip.GzipFile(path, "wb", mtime=0) as gz:
        gz.write(json.dumps(recs, separators=(",", ":")).encode("utf-8"))
    nh = sum(len(r["results"]) for r in recs)
    nr = sum(1 for r in recs if r["status"] != "ok")
    print(f"{name}: {len(recs)} messages, {nh} hits, {nr} raised, {path.stat().st_size} bytes")


def main():
    protocols = load_protocol_table()
    ref = ref_import.reference_class()()
    corp = Corpus(protocols)
    rng = random.Random(20261018)
    dump("reference_vectors.json.gz", run(ref, REFERENCE_VECTORS))
    dump("corpus_ms.json.gz", run(ref, [("MS", m) for m in batch_to_dicts(corp.pulse(pack.KIND_MS, 1500))]))
    dump("corpus_mu.json.gz", run(ref, [("MU", m) for m in batch_to_dicts(corp.pulse(pack.KIND_MU, 1500))]))
    dump("fuzz_ms.json.gz", run(ref, [("MS", m) for m in fuzz_pulse(rng, 1500, pack.KIND_MS, protocols)]))
    dump("fuzz_mu.json.gz", run(ref, [("MU", m) for m in fuzz_pulse(rng, 1500, pack.KIND_MU, protocols)]))
    # MC: "strict" = the reference as shipped, "repaired" = with the two documented one-line repairs
    mc = [("MC", m) for m in batch_to_dicts(corp.hexmsgs(pack.KIND_MC, 3000))]
    dump("corpus_mc_strict.json.gz", run(ref, mc))
    dump("corpus_mc_repaired.json.gz", run(ref_import.repaired_class()(), mc))
    dump("corpus_mn.json.gz", run(ref, [("MN", m) for m in batch_to_dicts(corp.hexmsgs(pack.KIND_MN, 3000))]))


if __name__ == "__main__":
    main()
