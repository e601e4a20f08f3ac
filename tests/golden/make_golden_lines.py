#!/usr/bin/env python3
"""Golden vectors for the line parser row (build container only): lines.json.gz.

Every line goes through the REFERENCE's SignalParser.parse_line (signalduino/parser/__init__.py:36-52) for rfmode None
and for one fixed rfmode; the decoded messages (protocol id, payload, metadata) and the RawFrame fields are recorded.
Lines: every firmware line literal of the reference's parser tests (ast: data, not code), corpus MS / MU / MC / MN
messages rendered as firmware lines, and seeded corruptions of all of them (framing, field order, duplicates,
empty parts, non-canonical values, unknown keys, lower-case types, reduced "Mred" payloads).
"""
import ast
import gzip
import json
import logging
import random
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))
from corpus.corpus import Corpus  # noqa: E402
from oracle import ref_import  # noqa: E402
from pysignalduino_b200 import pack  # noqa: E402
from pysignalduino_b200.protocol_data import load_protocol_table  # noqa: E402


def test_literals():
    out = []
    tests = ref_import.REFERENCE_ROOT / "tests"
    for f in ("test_ms_parser.py", "test_mu_parser.py", "test_mn_parser.py", "test_mc_parser.py", "test_decompress_payload.py",
              "test_mu_demodulation.py", "test_ms_demodulation.py", "test_controller.py"):
        p = tests / f
        if not p.exists():
            continue
        for node in ast.walk(ast.parse(p.read_text(encoding="utf-8"))):
            if isinstance(node, ast.Constant) and isinstance(node.value, str):
                v = node.value
                if 6 <= len(v) <= 2000 and (v.startswith("\x02M") or (v[:1] == "M" and v[2:3] == ";" and v.endswith(";"))):
                    out.append(v if v.startswith("\x02") else "\x02" + v + "\x03")
    return sorted(set(out))


def pulse_line(typ, d, rng):
    parts = [typ]
    pk = [k for k in d if k.startswith("P")]
    for k in pk:
        parts.append(f"{k}={d[k]}")
    tail = [f"D={d['data']}"]
    if typ == "MS":
        tail += [f"CP={d.get('CP', '0')}", f"SP={d.get('SP', '0')}"]
    else:
        tail += [f"CP={rng.randrange(8)}"]
    if "R" in d:
        tail.append(f"R={d['R']}")
    if typ == "MS" and rng.random() < 0.1:
        tail.append("O")
    if typ == "MS" and rng.random() < 0.1:
        tail.append(f"F={rng.randrange(256)}")
    if typ == "MU" and rng.random() < 0.2:
        tail.append(rng.choice(("O", "e", "p", f"w={rng.randrange(3)}")))
    if rng.random() < 0.3:
        rng.shuffle(tail)
    return "\x02" + ";".join(parts + tail) + ";\x03"


def corrupt(line, rng):
    body = line[1:-1]
    op = rng.randrange(16)
    if op == 0:
        return body                                               # no framing
    if op == 1:
        return line[:-1]                                          # ETX missing
    if op == 2:
        return "\x02" + body[0] + body[1].lower() + body[2:] + "\x03"   # lower-case type
    if op == 3:
        return "\x02" + body.replace(";D=", ";;D=", 1) + "\x03"   # empty part
    if op == 4:
        return "\x02" + body.replace("P1=", "P1=+", 1) + "\x03"   # value only float() understands
    if op == 5:
        return "\x02" + body.replace("P0=", "P00=", 1) + "\x03"   # multi-digit key, same id
    if op == 6:
        return "\x02" + body + "P1=77;" + "\x03"                  # duplicate pattern key after the tail
    if op == 7:
        i = body.find(";D=")
        return "\x02" + body[: i + 1] + "P1=77;" + body[i + 1 :] + "\x03" if i > 0 else line   # duplicate before D
    if op == 8:
        return "\x02" + body.replace("D=", "D=x", 1) + "\x03"     # non-digit in D
    if op == 9:
        return "\x02" + body.replace("CP=", "CP=a", 1) + "\x03"
    if op == 10:
        return "\x02" + body.replace(";R=", ";R=q", 1) + "\x03"
    if op == 11:
        return "\x02" + body + "zz=1;Q;" + "\x03"                 # unknown keys
    if op == 12:
        return "  " + line + "\r\n"                               # surrounding whitespace
    if op == 13:
        k = rng.randrange(3, max(4, len(body) - 1))
        return "\x02" + body[:k] + ";\x03"                         # truncated
    if op == 14:
        return "\x02" + body.replace("P2=", "P9=", 1).replace("P3=", "P8=", 1) + "\x03"
    i = rng.randrange(1, len(body))
    return "\x02" + body[:i] + rng.choice(";=-x07 ") + body[i + 1 :] + "\x03"


def reduce_line(line, rng):
    """A (valid) Mred=1 rendering of an MS / MU line: patterns as 3 bytes, D as one byte per two digits
    (base.py:113-127: high nibble = first digit, low 3 bits = second digit)."""
    body = line[1:-1].rstrip(";").split(";")
    typ, out = body[0], []
    for part in body[1:]:
        if part.startswith("P") and "=" in part and part[1:2].isdigit():
            idx, val = int(part[1]), int(part[3:])
            if idx > 7 or abs(val) > 32767:
                return None
            a = abs(val)
            lo, hi = a & 0xFF, a >> 8
            head = 0x80 | idx | (0x20 if val < 0 else 0) | (0x10 if lo & 0x80 else 0)
            out.append(chr(head) + chr((lo & 0x7F) | 0x80) + chr(hi | 0x80))
        elif part.startswith("D="):
            d = part[2:]
            if not d.isdigit() or any(c > "7" for c in d):
                return None
            key = "D"
            if len(d) % 2:
                d, key = d + "0", "d"
            out.append(key + "".join(chr((int(d[i]) << 4) | int(d[i + 1])) for i in range(0, len(d), 2)))
        elif part.startswith(("CP=", "SP=")) and len(part) == 4:
            out.append(part[0] + part[3])
        elif part.startswith("R=") and part[2:].isdigit() and int(part[2:]) < 256:
            out.append("R" + format(int(part[2:]), "X"))
        else:
            out.append(part)
    return "\x02" + typ + ";" + ";".join(out) + ";\x03"


def message_to_json():
    """MqttPublisher._message_to_json of the reference (signalduino/mqtt.py:228-245).  mqtt.py imports the MQTT client
    stack at module level (aiomqtt, paho, jsonschema: not installed here), so only that one function is compiled from
    the reference file (ast) and executed with the names it uses."""
    import json as _json
    from dataclasses import asdict

    types_mod = __import__("signalduino.types", fromlist=["RawFrame"])
    src = (ref_import.REFERENCE_ROOT / "signalduino" / "mqtt.py").read_text(encoding="utf-8")
    fn = next(n for n in ast.walk(ast.parse(src)) if isinstance(n, ast.FunctionDef) and n.name == "_message_to_json")
    fn.decorator_list = []
    ns = {"json": _json, "asdict": asdict, "RawFrame": types_mod.RawFrame, "DecodedMessage": types_mod.DecodedMessage}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "mqtt_message_to_json", "exec"), ns)
    return ns["_message_to_json"]


TO_JSON = None


def snapshot(msgs):
    out = []
    for m in msgs:
        f = m.raw
        out.append({"protocol_id": m.protocol_id, "payload": m.payload, "metadata": m.metadata, "json": TO_JSON(m),
                    "frame": {"line": f.line, "rssi": f.rssi, "freq_afc": f.freq_afc, "message_type": f.message_type}})
    return out


def main():
    global TO_JSON
    mod = ref_import.reference_parser_module()
    TO_JSON = message_to_json()
    log = logging.getLogger("golden_lines")
    log.addHandler(logging.NullHandler())
    log.propagate = False
    protocols = load_protocol_table()
    corp = Corpus(protocols)
    rng = random.Random(0x11E5)
    lines = test_literals()
    nlit = len(lines)
    for typ, kind, n in (("MS", pack.KIND_MS, 260), ("MU", pack.KIND_MU, 260)):
        b = corp.pulse(kind, n)
        for i in range(n):
            d = pack.unpack_pulse(b, i)
            if d.get("data"):
                lines.append(pulse_line(typ, d, rng))
    b = corp.hexmsgs(pack.KIND_MN, 160)
    for i in range(b.n):
        d = pack.unpack_hex(b, i)
        if d.get("data"):
            tail = (f"R={rng.randrange(256)};" if rng.random() < 0.7 else "") + (f"A={rng.randrange(-99, 100)};" if rng.random() < 0.5 else "")
            lines.append(f"\x02MN;D={'Y' if rng.random() < 0.1 else ''}{d['data']};{tail}\x03")
    b = corp.hexmsgs(pack.KIND_MC, 60)
    for i in range(b.n):
        d = pack.unpack_hex(b, i)
        if d.get("data"):
            lines.append(f"\x02MC;LL=-1017;LH=932;SL=-499;SH=486;D={d['data']};C={d.get('clock', 480)};L={d.get('bit_length', 64)};R={rng.randrange(256)};\x03")
    base = list(lines)
    for ln in base:
        for _ in range(2):
            lines.append(corrupt(ln, rng))
    for ln in base[nlit : nlit + 300]:
        r = reduce_line(ln, rng)
        if r:
            lines.append(r)
    # lines the tokenizer hands to the host path on purpose: more than 32 fields, longer than 1280 bytes, non-ASCII bytes
    for ln in base[nlit : nlit + 520 : 13]:
        body = ln[1:-1]
        lines.append("\x02" + body + "O;" * 40 + "\x03")
        lines.append("\x02" + body + "zz=" + "7" * 1400 + ";\x03")
        lines.append("\x02" + body + "q=\xe9\xb2;\x03")
        lines.append("\x02" + body.replace(";D=", ";" * 35 + "D=", 1) + "\x03")
    lines = list(dict.fromkeys(lines))
    recs = []
    nmsg = 0
    for rfmode in (None, "Bresser_5in1"):
        sp = mod.SignalParser(logger=log, rfmode=rfmode)
        for ln in lines:
            if rfmode is not None and "MN;" not in ln[:5].upper():
                continue
            res = snapshot(sp.parse_line(ln))
            nmsg += len(res)
            recs.append({"rfmode": rfmode, "line": ln, "payload": mod.base.extract_payload(ln), "results": res})
    with gzip.GzipFile(HERE / "lines.json.gz", "wb", mtime=0) as gz:
        gz.write(json.dumps(recs, separators=(",", ":")).encode("utf-8"))
    print(f"lines.json.gz: {len(lines)} lines, {len(recs)} parse_line calls, {nmsg} decoded messages, "
          f"{sum(1 for r in recs if r['results'])} lines with output, {(HERE / 'lines.json.gz').stat().st_size} bytes")


if __name__ == "__main__":
    main()
