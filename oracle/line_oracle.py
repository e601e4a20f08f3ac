"""CPU restatement of the reference's line parsers for MS / MU (TEST INFRASTRUCTURE ONLY — the product path is
csrc/sdb_lines.cu; only tests/ may import this).

    payload line -> parser dict, exactly as the reference builds it before calling SDProtocols.demodulate:
      MSParser.parse        signalduino/parser/ms.py:26-69   (_parse_to_dict :71-84, "D" required :41-46)
      MUParser.parse        signalduino/parser/mu.py:26-82   (validity regex :48-52, _parse_to_dict :84-95)
Pinned against the reference through tests/golden/lines.json.gz (tests/test_lines.py): the dict goes through
pack.pack_pulse and the C oracle, and the decoded messages must equal what the reference's SignalParser returned.
"""
from __future__ import annotations

import re
from typing import Any, Dict, Optional

MU_VALID = re.compile(r"^(?=.*D=\d+)(?:MU;(?:P[0-7]=-?[0-9]{1,5};){2,8}((?:D=\d{2,};)|(?:CP=\d;)|(?:R=\d+;)|(?:O;)|(?:e;)|(?:p;)|(?:w=\d;))*)$")


def parse_to_dict(line: str) -> Dict[str, Any]:
    d: Dict[str, Any] = {}
    for part in line.split(";"):
        if not part:
            continue
        if "=" in part:
            k, v = part.split("=", 1)
            d[k] = v
        else:
            d[part] = ""
    return d


def line_to_msg(payload: str, typ: str) -> Optional[Dict[str, Any]]:
    """The dict handed to demodulate(msg, typ), or None when the parser drops the line first."""
    if typ == "MU" and not MU_VALID.match(payload):
        return None
    d = parse_to_dict(payload)
    if "D" not in d:
        return None
    d["data"] = d["D"]
    return d
