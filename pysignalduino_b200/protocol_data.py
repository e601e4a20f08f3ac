"""Loader for the packaged protocol table (the data behind every decode decision).

Same content and iteration order as the reference's ``sd_protocols/protocols.json``
(loaded at sd_protocols/sd_protocols.py:30-41); re-serialised by
tools/import_protocol_table.py.
"""
from __future__ import annotations

import json
from pathlib import Path
from typing import Any, Dict

TABLE_PATH = Path(__file__).resolve().parent / "data" / "protocol_table.json"


def load_protocol_table() -> Dict[str, Dict[str, Any]]:
    """Return a fresh ``{protocol_id: {property: value}}`` dict in table order."""
    with open(TABLE_PATH, "r", encoding="utf-8") as f:
        return json.load(f)["protocols"]
