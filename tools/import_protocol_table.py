#!/usr/bin/env python3
"""Re-serialise the reference protocol table into the package's data file.

The table is DATA (160 protocol definitions), not code: the drop-in
``SDProtocols`` class must expose exactly the same property values (including
their JSON types: ``clockabs`` is a str in 14 entries, ``length_min`` a str in
157, ...) and iterate in exactly the same order as the reference
(``sd_protocols/protocols.json``, loaded at ``sd_protocols/sd_protocols.py:30-41``).

Run in the build container only (needs /root/reference):
    python tools/import_protocol_table.py
Writes ``pysignalduino_b200/data/protocol_table.json`` — one protocol per line,
compact separators, insertion order preserved.
"""
import json
import sys
from pathlib import Path

SRC = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference/sd_protocols/protocols.json")
DST = Path(__file__).resolve().parent.parent / "pysignalduino_b200" / "data" / "protocol_table.json"


def main() -> None:
    doc = json.loads(SRC.read_text(encoding="utf-8"))
    protos = doc["protocols"]
    lines = []
    for pid, props in protos.items():
        lines.append(json.dumps(pid) + ":" + json.dumps(props, separators=(",", ":"), ensure_ascii=False))
    body = "{\"version\":" + json.dumps(doc.get("version", "unknown")) + ",\n\"protocols\":{\n" + ",\n".join(lines) + "\n}}\n"
    # round-trip check: identical content and identical order
    back = json.loads(body)
    assert back["protocols"] == protos
    assert list(back["protocols"]) == list(protos)
    for pid in protos:
        assert list(back["protocols"][pid]) == list(protos[pid])
    DST.write_text(body, encoding="utf-8")
    print(f"wrote {DST} ({len(protos)} protocols, {len(body)} bytes)")


if __name__ == "__main__":
    main()
