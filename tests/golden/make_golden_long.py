#!/usr/bin/env python3
"""Golden vectors for long D streams (build container only): long_d.json.gz.

The reference puts no limit on len(D) (sd_protocols/message_unsynced.py:22-25; the MU parser regex is ``D=\\d{2,}``,
signalduino/parser/mu.py:48).  The fast kernels stage 1024 digits; longer messages (up to SDB_MAX_DIGITS = 4096) run
through the long kernels, and anything beyond is reported per message (status DomainError).  Every record here is decoded
by the REAL reference, twice: through ``SDProtocols.demodulate`` on the parser dict, and through
``SignalParser.parse_line`` on the firmware line.

Messages: corpus MS / MU rows (seeds 0x5D01 / 0x5D02) stretched to D lengths around every boundary —
1023, 1024, 1025 (fast / long kernel), 1500, 2000, 2047, 2048, 2049, 3000, 4095, 4096 (packed-domain limit), and 4097, 5000
(outside: expected status DomainError; the reference result is recorded so the test can show nothing else changed).
MU: the frame train is repeated (more repeats = more regex matches per protocol); MS: the valid frame, one chunk no
protocol knows (ends the chunk loop, message_synced.py:174-189), then repeats of the frame as filler.
"""
import gzip
import json
import logging
import random
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))
from corpus.corpus import Corpus, batch_to_dicts  # noqa: E402
from oracle import ref_import  # noqa: E402
from pysignalduino_b200 import pack  # noqa: E402
from pysignalduino_b200.protocol_data import load_protocol_table  # noqa: E402

LENGTHS = (1023, 1024, 1025, 1500, 2000, 2047, 2048, 2049, 3000, 4095, 4096, 4097, 5000)


def stretch(msg, typ, L, rng):
    m = dict(msg)
    d = m["data"]
    if typ == "MU":
        body = d
        # half of the MU messages get their repeats separated by a digit no symbol uses, so that runs end inside D
        sep = "" if rng.random() < 0.5 else next((c for c in "9876543210" if f"P{c}" not in m), "")
        out = body
        while len(out) < L:
            out += sep + body
        m["data"] = out[:L]
    else:
        unused = next((c for c in "9876543210" if f"P{c}" not in m), None)
        brk = (unused or d[-1]) * 2
        out = d + brk
        while len(out) < L:
            out += d
        m["data"] = out[:L]
    return m


def firmware_line(typ, m, rng):
    parts = [typ] + [f"{k}={v}" for k, v in m.items() if k.startswith("P")]
    tail = [f"D={m['data']}"]
    if typ == "MS":
        tail += [f"CP={m.get('CP', '0')}", f"SP={m.get('SP', '0')}"]
    else:
        tail += [f"CP={rng.randrange(8)}"]
    if "R" in m:
        tail.append(f"R={m['R']}")
    if rng.random() < 0.3:
        rng.shuffle(tail)
    return "\x02" + ";".join(parts + tail) + ";\x03"


def main():
    logging.disable(logging.CRITICAL)
    protocols = load_protocol_table()
    ref = ref_import.reference_class()()
    parser_mod = ref_import.reference_parser_module()
    parser = parser_mod.SignalParser(protocols=ref_import.reference_class()())
    corp = Corpus(protocols)
    rng = random.Random(0x10D6)
    recs = []
    for typ, kind in (("MS", pack.KIND_MS), ("MU", pack.KIND_MU)):
        base = [m for m in batch_to_dicts(corp.pulse(kind, 400)) if m.get("data") and len(m["data"]) >= 16]
        # keep rows the reference decodes (so that the stretched versions exercise accept paths) plus a few it rejects
        good = [m for m in base if ref_import.ref_demodulate(ref, m, typ)[1]]
        rest = [m for m in base if not ref_import.ref_demodulate(ref, m, typ)[1]]
        pick = good[:36] + rest[:6]
        for i, m0 in enumerate(pick):
            for L in LENGTHS:
                if (i + L) % 3 and L not in (1024, 1025, 4096, 4097):      # every row at the four boundaries, a third elsewhere
                    continue
                m = stretch(m0, typ, L, rng)
                st, res = ref_import.ref_demodulate(ref, m, typ)
                line = firmware_line(typ, m, rng)
                dec = parser.parse_line(line)
                recs.append({"type": typ, "msg": m, "dlen": len(m["data"]), "status": st, "results": [list(r) for r in res],
                             "line": line,
                             "line_results": [{"protocol_id": x.protocol_id, "payload": x.payload, "metadata": x.metadata} for x in dec]})
    path = HERE / "long_d.json.gz"
    with gzip.GzipFile(path, "wb", mtime=0) as gz:
        gz.write(json.dumps(recs, separators=(",", ":")).encode("utf-8"))
    by = {}
    for r in recs:
        k = (r["type"], r["dlen"])
        a = by.setdefault(k, [0, 0, 0])
        a[0] += 1
        a[1] += len(r["results"])
        a[2] += r["status"] != "ok"
    for k in sorted(by):
        print(k, "messages %d hits %d raised %d" % tuple(by[k]))
    print(f"long_d.json.gz: {len(recs)} records, {path.stat().st_size} bytes")


if __name__ == "__main__":
    main()
