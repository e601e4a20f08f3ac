"""Drop-in ``SDProtocols`` backed by the B200 batch demodulator.

Mirrors the public surface of the reference class (sd_protocols/sd_protocols.py:13-170): same
method names, same argument meaning, same returned dicts and the same exceptions.  The scalar
``demodulate(msg_data, msg_type)`` is a batch of one; ``demodulate_batch`` is the call the GPU
path exists for.  Every decode decision is taken on the device (libsdb200.so); this module only
packs inputs (pack.py), compiles the protocol table (table.py) and formats strings / dicts.
"""
from __future__ import annotations

import re
from collections.abc import Sequence
from typing import Any, Callable, Dict, List, Optional, Tuple

import numpy as np

from . import pack
from .capi import (HIT_LIST, HIT_MM_HOST, ST_DOMAIN, ST_OK, STATUS_EXC, STATUS_NAMES, Engine, Result)
from .protocol_data import load_protocol_table
from .table import HF_HAS_MAX, HF_HAS_MIN, PD_IDS, CompiledTable, compile_table
from .tracked import TrackedDict, Version

# SDB_M_* ids of the manchester.py decoders (csrc/sdb_table.h)
MC_METHODS = {"mcBit2Funkbus": 1, "mcBit2Sainlogic": 2, "mcBit2AS": 3, "mcBit2Hideki": 4, "mcBit2Maverick": 5,
              "mcBit2OSV1": 6, "mcBit2OSV2o3": 7, "mcBit2OSPIR": 8, "mcRaw": 9, "mcBit2TFA": 11, "mcBit2Grothe": 12,
              "mcBit2SomfyRTS": 13}


_STATUS_ARRAY = np.array([STATUS_NAMES.get(i, f"status{i}") for i in range(256)], dtype=object)


class BatchResults(Sequence):
    """``results`` of :meth:`SDProtocols.demodulate_batch`: ``results[i]`` is the list of hit dicts of message i, exactly what
    the reference's ``demodulate()`` returns for it, materialised on access from the raw result arrays (27 M hits per 10 M
    messages are not turned into Python objects unless somebody looks at them).  Compares equal to the equivalent list."""

    def __init__(self, sdp: "SDProtocols", batch, res: Result, pool: bytes):
        self._sdp, self._batch, self._res, self._pool = sdp, batch, res, pool       # res.hits: pack.PAYHIT_DTYPE records
        self._ids = sdp.engine().table.ids
        self._clk: Dict[int, float] = {}

    def __len__(self) -> int:
        return self._batch.n

    def _payload(self, i: int) -> str:
        a = int(self._res.hits["off"][i])
        return self._pool[a : self._pool.index(0, a)].decode("latin-1")

    def __getitem__(self, m):
        if isinstance(m, slice):
            return [self[i] for i in range(*m.indices(len(self)))]
        if m < 0:
            m += len(self)
        if not 0 <= m < len(self):
            raise IndexError("message index out of range")
        res, batch, kind, ids = self._res, self._batch, self._batch.kind, self._ids
        o = res.out[m]
        if int(o["status"]) != ST_OK:
            return []
        lst: List[Dict[str, Any]] = []
        h0 = int(o["hit_off"])
        for i in range(h0, h0 + int(o["nhits"])):
            h = res.hits[i]
            fl = int(h["flags"])
            if (fl & HIT_LIST) and int(h["aux"]) != 0:
                continue
            pi = int(h["proto"])
            payload = self._payload(i)
            if fl & HIT_MM_HOST:                                      # a modulematch shape the device program cannot express
                if not re.search(str(self._sdp._protocols[ids[pi]].get("modulematch")), payload):       # message_unsynced.py:277-280
                    continue
            if kind == pack.KIND_MS:
                lst.append({"protocol_id": ids[pi], "payload": payload,
                            "meta": {"bit_length": int(h["nbits"]), "rssi": batch.rssi[m], "clock": float(batch.clock[m])}})   # message_synced.py:239
            elif kind == pack.KIND_MU:
                if pi not in self._clk:
                    self._clk[pi] = float(self._sdp.check_property(ids[pi], "clockabs", 1))
                lst.append({"protocol_id": ids[pi], "payload": payload,
                            "meta": {"bit_length": int(h["nbits"]), "rssi": batch.rssi[m], "clock": self._clk[pi]}})          # message_unsynced.py:288
            elif kind == pack.KIND_MC:
                pid = batch.protocol_ids[m]
                lst.append({"protocol_id": str(pid), "payload": payload,
                            "meta": {"protocol_id": pid, "rssi": None, "freq_afc": None}})                                   # manchester.py:136-140
            else:
                meta = {"is_raw": False} if int(h["aux"]) in (18, 19, 20) else {}                                            # helpers.py:578,627,715
                lst.append({"protocol_id": batch.protocol_ids[m], "payload": payload, "meta": meta})
        return lst

    def __eq__(self, other):
        if isinstance(other, (list, tuple, BatchResults)):
            return len(other) == len(self) and all(a == b for a, b in zip(self, other))
        return NotImplemented

    def __repr__(self) -> str:
        return f"<BatchResults: {len(self)} messages, {len(self._res.hits)} hits>"


class SDProtocols:
    """Protocol handling class: same API as the reference, demodulation on the GPU.

    Extra keyword arguments (not in the reference):
      device       CUDA device index of this instance's engine (default 0)
      mc_repaired  False = MC path exactly as shipped (raises TypeError, SURVEY §8c "strict");
                   True  = with the two documented one-line repairs of manchester.py:83/:120.
    """

    def __init__(self, device: int = 0, mc_repaired: bool = False):
        self._ver = Version()
        self._table: Optional[CompiledTable] = None
        self._table_ver = -1
        self._engine: Optional[Engine] = None
        self._engine_ver = -1
        self._log_callback: Optional[Callable[[str, int], None]] = None
        self._protocols = self._load_protocols()
        self.set_defaults()
        self._device = device
        self.mc_repaired = mc_repaired

    # The protocol dict is wrapped in mutation-tracking containers (tracked.py): callers may edit rows between calls as they
    # can with the reference (tests/test_manchester_protocols.py:54,84), and the compiled device table follows.
    @property
    def _protocols(self) -> Dict[str, Any]:
        return self._proto_dict

    @_protocols.setter
    def _protocols(self, value: Dict[str, Any]) -> None:
        self._proto_dict = TrackedDict(value, self._ver)
        self._ver.n += 1

    # ------------------------------------------------------------------ table / property API
    def _load_protocols(self) -> Dict[str, Any]:
        """sd_protocols.py:30-41"""
        return load_protocol_table()

    def protocol_exists(self, pid: str) -> bool:
        return pid in self._protocols

    def get_protocol_list(self) -> dict:
        return self._protocols

    def get_keys(self, filter_key: str = None) -> list:
        if filter_key:
            return [pid for pid, props in self._protocols.items() if filter_key in props]
        return list(self._protocols.keys())

    def check_property(self, pid: str, value_name: str, default=None):
        return self._protocols.get(pid, {}).get(value_name, default)

    def get_property(self, pid: str, value_name: str):
        return self._protocols.get(pid, {}).get(value_name)

    def set_defaults(self):
        """sd_protocols.py:157-160"""
        for pid, proto in self._protocols.items():
            proto.setdefault("active", True)
            proto.setdefault("name", f"Protocol_{pid}")

    def register_log_callback(self, callback):
        if callable(callback):
            self._log_callback = callback

    def _logging(self, message: str, level: int = 3):
        if self._log_callback:
            self._log_callback(message, level)

    # ------------------------------------------------------------------ engine management
    def compiled_table(self) -> CompiledTable:
        """The compiled form of the CURRENT protocol dict (host only, no GPU needed)."""
        v = self._ver.n
        if self._table is None or v != self._table_ver:
            self._table = compile_table(self._protocols)
            self._table_ver = v
            for pid, why in self._table.unsupported.items():
                self._logging(f"protocol {pid} is not compiled into the device table and never matches: {why}", 2)
        return self._table

    def engine(self) -> Engine:
        """The device engine for the CURRENT protocol dict (recompiled when the dict was mutated: one integer compare
        per call, see tracked.py)."""
        v = self._ver.n
        if self._engine is None or v != self._engine_ver:
            if self._engine is not None:
                self._engine.close()
                self._engine = None
            self._engine = Engine(self.compiled_table(), self._device)
            self._engine_ver = v
        return self._engine

    # ------------------------------------------------------------------ batch API (new)
    def demodulate_packed(self, batch) -> Result:
        """Run a packed batch (pack.PulseBatch / pack.HexBatch) and return the raw result arrays."""
        return self.engine().demod_host(batch, mc_repaired=self.mc_repaired)

    def pack(self, msgs: Sequence[Dict[str, Any]], msg_type: str):
        kind = pack.KIND_BY_NAME[msg_type]
        if kind in (pack.KIND_MS, pack.KIND_MU):
            return pack.pack_pulse(msgs, kind)
        eng = self.engine()
        index = {pid: i for i, pid in enumerate(eng.table.ids)}
        return pack.pack_hex(msgs, kind, index)

    def demodulate_batch(self, msgs: Sequence[Dict[str, Any]], msg_type: str, lazy: bool = True):
        """Demodulate many messages of one type.

        Returns (statuses, results): statuses[i] is "ok" or the name of the exception the reference
        raises for message i ("IndexError", "TypeError", "ValueError"); results[i] is the list the
        reference's demodulate() returns (empty when it raises).  A message the packed domain cannot
        represent (pack.py) gets status "DomainError" and [] — it was NOT decoded; every other message
        of the batch is.

        The dicts are packed natively (csrc/sdb_fastpack.c), decoded and formatted on the device
        (sdb_demod_host_payloads); ``results`` is a :class:`BatchResults` sequence that builds the per-message
        lists on access (``lazy=False`` returns plain lists).
        """
        if msg_type not in pack.KIND_BY_NAME:
            self._logging(f"Unknown message type {msg_type}", 3)
            return ["ok"] * len(msgs), [[] for _ in msgs]
        batch = self.pack(msgs, msg_type)
        res, pool = self.engine().demod_payloads(batch, mc_repaired=self.mc_repaired)
        statuses = _STATUS_ARRAY[res.out["status"]].tolist()
        results = BatchResults(self, batch, res, pool.tobytes())
        return statuses, (results if lazy else list(results))

    def format_results(self, batch, res: Result, msgs: Optional[Sequence[Dict[str, Any]]] = None):
        eng = self.engine()
        ids = eng.table.ids
        kind = batch.kind
        pool, off = eng.format_hits(kind, res.hits, res.bits)
        text = pool.decode("latin-1")
        statuses = [STATUS_NAMES[int(s)] for s in res.out["status"]]
        results: List[List[Dict[str, Any]]] = []
        hits = res.hits
        clk_cache: Dict[int, float] = {}
        for m in range(batch.n):
            o = res.out[m]
            lst: List[Dict[str, Any]] = []
            h0 = int(o["hit_off"])
            for i in range(h0, h0 + int(o["nhits"])):
                h = hits[i]
                if (int(h["flags"]) & HIT_LIST) and int(h["aux"]) != 0:
                    continue
                pi = int(h["proto"])
                payload = text[int(off[i]) : int(off[i + 1])]
                if int(h["flags"]) & HIT_MM_HOST:                          # a modulematch shape the device program cannot express
                    if not re.search(str(self._protocols[ids[pi]].get("modulematch")), payload):     # message_unsynced.py:277-280
                        continue
                if kind in (pack.KIND_MS, pack.KIND_MU):
                    if kind == pack.KIND_MS:
                        clock = float(batch.clock[m])                      # message_synced.py:239
                    else:
                        if pi not in clk_cache:
                            clk_cache[pi] = float(self.check_property(ids[pi], "clockabs", 1))
                        clock = clk_cache[pi]                              # message_unsynced.py:288
                    lst.append({"protocol_id": ids[pi], "payload": payload,
                                "meta": {"bit_length": int(h["nbits"]), "rssi": batch.rssi[m], "clock": clock}})
                elif kind == pack.KIND_MC:
                    pid = batch.protocol_ids[m]
                    lst.append({"protocol_id": str(pid), "payload": payload,
                                "meta": {"protocol_id": pid, "rssi": None, "freq_afc": None}})   # manchester.py:136-140
                else:
                    meta = {"is_raw": False} if int(h["aux"]) in (18, 19, 20) else {}             # helpers.py:578,627,715
                    lst.append({"protocol_id": batch.protocol_ids[m], "payload": payload, "meta": meta})
            results.append(lst)
        return statuses, results

    # ------------------------------------------------------------------ reference API
    def demodulate(self, msg_data: Dict[str, Any], msg_type: str) -> list:
        """sd_protocols.py:60-74"""
        if msg_type == "MS":
            return self.demodulate_ms(msg_data, msg_type)
        elif msg_type == "MC":
            return self.demodulate_mc(msg_data, msg_type)
        elif msg_type == "MN":
            return self.demodulate_mn(msg_data, msg_type)
        elif msg_type == "MU":
            return self.demodulate_mu(msg_data, msg_type)
        self._logging(f"Unknown message type {msg_type}", 3)
        return []

    def _raise_status(self, status: int, batch, what: str):
        if status == ST_DOMAIN:
            raise pack.DomainError(f"{what}: outside the packed domain, not decoded: {batch.domain.get(0, 'malformed record')}")
        raise STATUS_EXC[status](f"reference {what} raises {STATUS_NAMES[status]} on this message")

    def _one(self, msg_data: Dict[str, Any], msg_type: str) -> list:
        batch = self.pack([msg_data], msg_type)
        res, pool = self.engine().demod_payloads(batch, mc_repaired=self.mc_repaired)
        st = int(res.out["status"][0])
        if st != ST_OK:
            self._raise_status(st, batch, f"demodulate_{msg_type.lower()}")
        return BatchResults(self, batch, res, pool.tobytes())[0]

    def demodulate_ms(self, msg_data: Dict[str, Any], msg_type: str = "MS") -> List[Dict[str, Any]]:
        """message_synced.py:10-243 (invalid input -> [] with a level-3 log line, :21-47)"""
        raw_data = msg_data.get("data", "")
        if not raw_data or not raw_data.isdigit():
            self._logging(f"MS Demod: Invalid rawData D=: {raw_data}", 3)
            return []
        for key, label in (("CP", "CP"), ("SP", "SP")):
            v = msg_data.get(key, "")
            if not v or not v.isdigit():
                self._logging(f"MS Demod: Invalid {label}: {v}", 3)
                return []
        if "R" in msg_data and not msg_data.get("R", "").isdigit():
            self._logging(f"MS Demod: Invalid RSSI R=: {msg_data.get('R', '')}", 3)
            return []
        return self._one(msg_data, "MS")

    def demodulate_mu(self, msg_data: Dict[str, Any], msg_type: str = "MU") -> List[Dict[str, Any]]:
        """message_unsynced.py:11-296"""
        raw_data = msg_data.get("data", "")
        if not raw_data:
            self._logging(f"MU Demod: Invalid rawData D=: {raw_data}", 3)
            return []
        return self._one(msg_data, "MU")

    def demodulate_mc(self, msg_data: Dict[str, Any], msg_type: str, version: str | None = None) -> list:
        """sd_protocols.py:76-111"""
        protocol_id = msg_data.get("protocol_id")
        if not protocol_id or not self.protocol_exists(protocol_id):
            self._logging(f"MC Demodulation failed: Protocol ID {protocol_id} not found or missing.", 3)
            return []
        toggle = msg_type == "Mc" or bool(version and version[:6] == "V 3.2.")      # manchester.py:94
        eng = self.engine()
        index = {pid: i for i, pid in enumerate(eng.table.ids)}
        batch = pack.pack_hex([msg_data], pack.KIND_MC, index, toggle_polarity=toggle)
        res = eng.demod_host(batch, mc_repaired=self.mc_repaired)
        st = int(res.out["status"][0])
        if st != ST_OK:
            self._raise_status(st, batch, "demodulate_mc")
        return self.format_results(batch, res)[1][0]

    def demodulate_mn(self, msg_data: Dict[str, Any], msg_type: str) -> list:
        """sd_protocols.py:113-155"""
        if "protocol_id" not in msg_data:
            self._logging(f"MN Demodulation failed: Missing protocol_id in msg_data: {msg_data}", 3)
            return []
        protocol_id = msg_data["protocol_id"]
        if not self.protocol_exists(protocol_id):
            self._logging(f"MN Demodulation: Protocol ID {protocol_id} not found.", 3)
            return []
        if not self.get_property(protocol_id, "method"):
            self._logging(f"MN Demodulation: No method defined for protocol {protocol_id}. Data: {msg_data.get('data', '')}", 3)
            return []
        return self._one(msg_data, "MN")

    # ------------------------------------------------------------------ Conv* (MN converters, batch of one on the device)
    def _conv(self, method: int, msg_data: Dict[str, Any]) -> list:
        eng = self.engine()
        batch = pack.pack_hex([msg_data], pack.KIND_MN, {}, method_override=method)
        res = eng.demod_host(batch, mc_repaired=self.mc_repaired)
        st = int(res.out["status"][0])
        if st != ST_OK:
            self._raise_status(st, batch, "MN converter")
        return self.format_results(batch, res)[1][0]

    def ConvBresser_lightning(self, msg_data, msg_type="MN"):
        """helpers.py:223-279"""
        return self._conv(14, msg_data)

    def ConvBresser_5in1(self, msg_data, msg_type="MN"):
        """helpers.py:382-425"""
        return self._conv(15, msg_data)

    def ConvBresser_6in1(self, msg_data, msg_type="MN"):
        """helpers.py:427-471"""
        return self._conv(16, msg_data)

    def ConvBresser_7in1(self, msg_data, msg_type="MN"):
        """helpers.py:473-523"""
        return self._conv(17, msg_data)

    def ConvPCA301(self, msg_data, msg_type="MN"):
        """helpers.py:525-579"""
        return self._conv(18, msg_data)

    def ConvKoppFreeControl(self, msg_data, msg_type="MN"):
        """helpers.py:581-628"""
        return self._conv(19, msg_data)

    def ConvLaCrosse(self, msg_data, msg_type="MN"):
        """helpers.py:630-716"""
        return self._conv(20, msg_data)

    # ------------------------------------------------------------------ postDemo_* (unit ops on the device)
    def _postdemo(self, method: str, bit_msg_array):
        rc, out = self.engine().unit_postdemod(PD_IDS[method], bit_msg_array)
        if rc == -2:
            raise ValueError("invalid literal for int() with base 2: ''")    # postdemodulation.py:471
        if rc < 1:
            return (0, None)
        return (1, out)

    def postDemo_EM(self, name, bit_msg_array):
        """postdemodulation.py:27-88"""
        return self._postdemo("postDemo_EM", bit_msg_array)

    def postDemo_Revolt(self, name, bit_msg_array):
        """postdemodulation.py:90-137"""
        return self._postdemo("postDemo_Revolt", bit_msg_array)

    def postDemo_FS20(self, name, bit_msg_array):
        """postdemodulation.py:139-243"""
        return self._postdemo("postDemo_FS20", bit_msg_array)

    def postDemo_FHT80(self, name, bit_msg_array):
        """postdemodulation.py:245-337"""
        return self._postdemo("postDemo_FHT80", bit_msg_array)

    def postDemo_FHT80TF(self, name, bit_msg_array):
        """postdemodulation.py:339-423"""
        return self._postdemo("postDemo_FHT80TF", bit_msg_array)

    def postDemo_WS2000(self, name, bit_msg_array):
        """postdemodulation.py:425-578"""
        return self._postdemo("postDemo_WS2000", bit_msg_array)

    def postDemo_WS7035(self, name, bit_msg_array):
        """postdemodulation.py:580-640"""
        return self._postdemo("postDemo_WS7035", bit_msg_array)

    def postDemo_WS7053(self, name, bit_msg_array):
        """postdemodulation.py:642-706"""
        return self._postdemo("postDemo_WS7053", bit_msg_array)

    def postDemo_lengtnPrefix(self, name, bit_msg_array):
        """postdemodulation.py:708-730"""
        return self._postdemo("postDemo_lengtnPrefix", bit_msg_array)

    # ------------------------------------------------------------------ mcBit2* / mcRaw (unit ops on the device)
    _MC_REASONS = {1: "message is too short", 2: "message is too long", 3: "wrong bits at begin", 4: "parity error",
                   5: "checksum error", 9: "sync not found"}

    def _mc_unit(self, method: int, name, bit_data, protocol_id, mcbitnum):
        """One manchester.py decoder call through sdb_unit_mc; the host only renders hex and message strings."""
        if mcbitnum is None:
            mcbitnum = len(bit_data)
        if any(c not in "01" for c in bit_data):
            raise pack.DomainError("mcBit2*: bit_data must be a string of 0/1")
        eng = self.engine()
        try:
            index = eng.table.ids.index(protocol_id)
        except ValueError:
            index = 0xFFFFFFFF                                   # check_property falls back to its defaults
        try:
            is119 = int(protocol_id) == 119                      # manchester.py:242
        except (TypeError, ValueError):
            is119 = False
        if index != 0xFFFFFFFF:
            is119 = False                                        # the table row carries it
        rc, reason, parts = eng.unit_mc(index, bit_data, int(mcbitnum), method | (0x100 if is119 else 0))
        if rc <= -100:
            raise STATUS_EXC[-rc - 100](f"reference raises in the MC decoder ({method})")
        if rc == 1:
            if method == MC_METHODS["mcBit2TFA"]:
                return (1, [self.bin_str_2_hex_str(x) for x in parts])
            return (1, self.bin_str_2_hex_str(parts[0]))
        base = reason & 0xFF
        if base in self._MC_REASONS:
            return (-1, self._MC_REASONS[base])
        if base == 6:
            return (-1, f"{name}: lib/mcBit2Sainlogic, start 010100 not found")
        if base == 7:
            return (-1, f"message must be 32 bits, got {mcbitnum}")
        if base == 8:
            n = len(bit_data[1:57]) if mcbitnum == 57 else len(bit_data)         # manchester.py:783-789
            return (-1, f"message must be 56 bits, got {n}")
        if base == 10:
            tail = {0: "", 0x100: ", message is too short", 0x200: ", message is too long",
                    0x400: ", protocol does not exists"}[reason & 0x700]
            return (-1, f" no duplicate found{tail}")
        if base == 11:
            return (-1, f"loop error, please report this data {bit_data}")
        raise RuntimeError(f"sdb_unit_mc: unknown reason code {reason}")

    def mcBit2Funkbus(self, name, bit_data, protocol_id, mcbitnum=None):
        """manchester.py:207-300"""
        return self._mc_unit(MC_METHODS["mcBit2Funkbus"], name, bit_data, protocol_id, mcbitnum)

    def mcBit2Sainlogic(self, name, bit_data, protocol_id, mcbitnum=None):
        """manchester.py:302-354"""
        return self._mc_unit(MC_METHODS["mcBit2Sainlogic"], name, bit_data, protocol_id, mcbitnum)

    def mcBit2AS(self, name, bit_data, protocol_id, mcbitnum=None):
        """manchester.py:356-416"""
        return self._mc_unit(MC_METHODS["mcBit2AS"], name, bit_data, protocol_id, mcbitnum)

    def mcBit2Hideki(self, name, bit_data, protocol_id, mcbitnum=None):
        """manchester.py:418-450"""
        return self._mc_unit(MC_METHODS["mcBit2Hideki"], name, bit_data, protocol_id, mcbitnum)

    def mcBit2Maverick(self, name, bit_data, protocol_id, mcbitnum=None):
        """manchester.py:452-484"""
        return self._mc_unit(MC_METHODS["mcBit2Maverick"], name, bit_data, protocol_id, mcbitnum)

    def mcBit2OSV1(self, name, bit_data, protocol_id, mcbitnum=None):
        """manchester.py:486-518"""
        return self._mc_unit(MC_METHODS["mcBit2OSV1"], name, bit_data, protocol_id, mcbitnum)

    def mcBit2OSV2o3(self, name, bit_data, protocol_id, mcbitnum=None):
        """manchester.py:520-552"""
        return self._mc_unit(MC_METHODS["mcBit2OSV2o3"], name, bit_data, protocol_id, mcbitnum)

    def mcBit2OSPIR(self, name, bit_data, protocol_id, mcbitnum=None):
        """manchester.py:554-586"""
        return self._mc_unit(MC_METHODS["mcBit2OSPIR"], name, bit_data, protocol_id, mcbitnum)

    def mcRaw(self, name, bit_data, protocol_id, mcbitnum, other_arg=None):
        """manchester.py:588-613 (the ManchesterMixin definition wins over helpers.mcraw in the MRO)"""
        return self._mc_unit(MC_METHODS["mcRaw"], name, bit_data, protocol_id, int(mcbitnum))    # :608 int(None) raises

    def mcraw(self, name="anonymous", bit_data=None, protocol_id=None, mcbitnum=None):
        """helpers.py:90-122 (lower case: the rfmode-less '57' table entry names it; note its own message text)"""
        if bit_data is None:
            return (-1, "no bitData provided")
        if protocol_id is None:
            return (-1, "no protocolId provided")
        if isinstance(bit_data, str) and bit_data and any(c not in "01" for c in bit_data):
            # the length rule is checked first (:112-115), then bin_str_2_hex_str rejects the string (:118-120)
            rc, msg = self._mc_unit(10, name, "0" * len(bit_data), protocol_id, len(bit_data) if mcbitnum is None else mcbitnum)
            return (-1, "message is to long") if rc == -1 else (-1, "invalid bit data")
        rc, msg = self._mc_unit(10, name, bit_data, protocol_id, mcbitnum)
        return (rc, "message is to long") if rc == -1 else (rc, msg)

    def mcBit2TFA(self, name, bit_data, protocol_id, mcbitnum=None):
        """manchester.py:615-719 (returns the LIST of duplicated parts)"""
        return self._mc_unit(MC_METHODS["mcBit2TFA"], name, bit_data, protocol_id, mcbitnum)

    def mcBit2Grothe(self, name, bit_data, protocol_id, mcbitnum=None):
        """manchester.py:721-754"""
        return self._mc_unit(MC_METHODS["mcBit2Grothe"], name, bit_data, protocol_id, mcbitnum)

    def mcBit2SomfyRTS(self, name, bit_data, protocol_id, mcbitnum=None):
        """manchester.py:756-795"""
        return self._mc_unit(MC_METHODS["mcBit2SomfyRTS"], name, bit_data, protocol_id, mcbitnum)

    # ------------------------------------------------------------------ MC / MN internals reachable by name (manchester.py)
    def _mc_reason_text(self, reason: int, name: str, method_name_full, bit_len: int, bit_data: Optional[str]) -> str:
        base = reason & 0xFF
        if base in self._MC_REASONS:
            return self._MC_REASONS[base]
        if base == 6:
            return f"{name}: lib/mcBit2Sainlogic, start 010100 not found"
        if base == 7:
            return f"message must be 32 bits, got {bit_len}"
        if base == 8:
            return f"message must be 56 bits, got {bit_len - 1 if bit_len == 57 else bit_len}"      # manchester.py:783-789
        if base in (10, 12, 13, 14):
            return " no duplicate found" + {10: "", 12: ", message is too short", 13: ", message is too long",
                                            14: ", protocol does not exists"}[base]
        if base == 11:
            return f"loop error, please report this data {bit_data}"
        if base == 20:
            return "clock out of range"
        if base == 22:
            return f"Unknown protocol method {method_name_full}"
        raise RuntimeError(f"unknown MC reason code {reason}")

    def _convert_mc_hex_to_bits(self, name, raw_hex, polarity_invert, hlen):
        """manchester.py:18-47: optional (upper-case only) nibble inversion, then hex_to_bin_str.  Pure string formatting:
        the batch path does the same with index arithmetic on the nibble stream (McBits in csrc/sdb_hex.cu)."""
        text = raw_hex
        if polarity_invert:
            text = "".join("FEDCBA9876543210"["0123456789ABCDEF".index(c)] if c in "0123456789ABCDEF" else c for c in raw_hex)
        bit_data = self.hex_to_bin_str(text)
        self._logging(f"{name}: extracted data {bit_data} (bin)", 5)
        return (1, bit_data)

    def _demodulate_mc_data(self, name, protocol_id, clock, raw_hex, mcbitnum, messagetype, version):
        """manchester.py:49-144 as one device call: the checks, the decoder and the reject reason all come from the MC kernel
        (``SdbMsgOut.reason``); the host renders the reference's message strings."""
        eng = self.engine()
        pid = protocol_id
        if not isinstance(pid, str) or pid not in self._protocols:
            # check_property falls back to its defaults for an unknown id: limits -1 / 9999, no clockrange, no method
            if mcbitnum < -1:
                return (-1, "message is too short", {})
            if mcbitnum > 9999:
                return (-1, "message is too long", {})
            return [(-1, "Protocol method not defined", {})]
        toggle = messagetype == "Mc" or bool(version and version[:6] == "V 3.2.")      # manchester.py:94
        msg = {"protocol_id": pid, "data": raw_hex, "clock": clock, "bit_length": mcbitnum}
        batch = pack.pack_hex([msg], pack.KIND_MC, {pid: eng.table.ids.index(pid)}, toggle_polarity=toggle)
        res = eng.demod_host(batch, mc_repaired=self.mc_repaired)
        st, reason = int(res.out["status"][0]), int(res.out["reason"][0])
        if reason == 21:
            return [(-1, "Protocol method not defined", {})]                           # :108-109 (a list, as shipped)
        if st != ST_OK:
            self._raise_status(st, batch, "_demodulate_mc_data")
        if int(res.out["nhits"][0]) == 0:
            # the decoders' messages quote len(bit_data) / bit_data: the string the device walked (inverted, leading zero nibbles dropped)
            inv = (self.check_property(pid, "polarity", "") == "invert") ^ toggle
            bits = self._convert_mc_hex_to_bits(name, raw_hex, inv, len(raw_hex))[1] or ""
            return (-1, self._mc_reason_text(reason, name, self.get_property(pid, "method"), len(bits), bits), {})
        pool, off = eng.format_hits(pack.KIND_MC, res.hits, res.bits)
        dmsg = pool[int(off[0]) : int(off[1])].decode("latin-1")
        return (1, dmsg, {"protocol_id": protocol_id, "rssi": None, "freq_afc": None})        # :134-142

    def _demodulate_mn_data(self, name, protocol_id, msg_data):
        """manchester.py:147-204: method lookup, one converter call (on the device), first well-formed result."""
        method_name_full = self.get_property(protocol_id, "method")
        if not method_name_full:
            return []
        method_name = method_name_full.split(".")[-1]
        func = getattr(self, method_name, None)
        if not callable(func):
            return []
        try:
            demodulated_list = func(msg_data, "MN")
        except TypeError:
            return []
        if not isinstance(demodulated_list, list) or not demodulated_list:
            return []
        for decoded in demodulated_list:
            if not isinstance(decoded, dict) or "protocol_id" not in decoded:
                continue
            return [{"protocol_id": str(decoded["protocol_id"]), "payload": str(decoded.get("payload", "")),
                     "meta": decoded.get("meta", {})}]
        return []

    # ------------------------------------------------------------------ checksum helpers (helpers.py), scalar API surface
    # The batch path computes these inside the MN kernel (lfsr16 / crc16 / the LaCrosse CRC in csrc/sdb_hex.cu); the methods
    # exist because the reference's converters and tests call them by name.
    def lfsr_digest16(self, bytes_count, gen, key, raw_data):
        """helpers.py:190-221: Galois LFSR digest over the first bytes_count bytes of a hex string (0 when it is too short
        or not hex)."""
        if len(raw_data) < 2 * bytes_count:
            return 0
        try:
            data = [int(raw_data[2 * k : 2 * k + 2], 16) for k in range(bytes_count)]
        except ValueError:
            return 0
        digest = 0
        for byte in data:
            for bit in (0x80, 0x40, 0x20, 0x10, 0x08, 0x04, 0x02, 0x01):
                if byte & bit:
                    digest ^= key
                key = (key >> 1) ^ gen if key & 1 else key >> 1
        return digest

    def _calc_crc16(self, hex_data, poly, init, refin, refout, xorout):
        """helpers.py:281-309: bitwise CRC-16, MSB first, optional input / output reflection; '0000' for non-hex input."""
        try:
            data = bytes.fromhex(hex_data)
        except ValueError:
            self._logging(f"_calc_crc16: Invalid hex data provided: {hex_data}", 3)
            return "0000"
        reg = init
        for byte in data:
            if refin:
                byte = int(format(byte, "08b")[::-1], 2)
            reg ^= byte << 8
            for _ in range(8):
                reg = ((reg << 1) ^ poly if reg & 0x8000 else reg << 1) & 0xFFFF
        if refout:
            reg = int(format(reg, "016b")[::-1], 2)
        return format(reg ^ xorout, "04X")

    def _calc_crc8_la_crosse(self, hex_data):
        """helpers.py:311-380: what the method returns is its last loop — a right-shifting CRC-8 with 0x31 XORed in on a set
        low bit (bytes.fromhex raises ValueError on bad input, as there)."""
        reg = 0
        for byte in bytes.fromhex(hex_data):
            reg ^= byte
            for _ in range(8):
                reg = ((reg >> 1) ^ 0x31 if reg & 1 else reg >> 1) & 0xFF
        return reg

    # ------------------------------------------------------------------ RSL placeholders (rsl_handler.py:12-55)
    def decode_rsl(self, bit_data):
        """rsl_handler.py:12-33 (a placeholder upstream as well: echoes its input)"""
        self._logging(f"lib/decode_rsl, bit_data length: {len(str(bit_data))}", 5)
        return {"decoded": str(bit_data), "status": 1}

    def encode_rsl(self, data):
        """rsl_handler.py:35-55"""
        self._logging(f"lib/encode_rsl, data: {data}", 5)
        return {"encoded": str(data), "status": 1}

    # ------------------------------------------------------------------ small pure helpers (API surface)
    def bin_str_2_hex_str(self, num):
        """helpers.py:28-64: right-aligned nibbles, upper case; '' -> ''; non-binary -> None."""
        if num is None:
            return None
        if not num:
            return ""
        if not isinstance(num, str) or any(c not in "01" for c in num):
            return None
        ndig = (len(num) + 3) // 4
        return format(int(num, 2), "X").zfill(ndig)

    def hex_to_bin_str(self, hex_string):
        """helpers.py:168-188: leading zero nibbles are dropped ('00FF' -> '11111111')."""
        if hex_string is None:
            return None
        try:
            b = bin(int(hex_string, 16))[2:]
        except ValueError:
            return None
        return b.zfill((len(b) + 3) // 4 * 4)

    def length_in_range(self, protocol_id, message_length):
        """helpers.py:124-166 — answered from the compiled table's per-protocol length rules (the same row the MC / MN
        kernels and ``in_range`` in csrc/sdb_hex.cu read)."""
        tab = self.compiled_table()
        pid = str(protocol_id)
        if pid not in self._protocols:
            return (0, "protocol does not exists")
        row = tab.hex_rows[tab.ids.index(pid)]
        if protocol_id in self._protocols:                      # check_property / get_property look the id up as passed (int ids miss)
            fl = int(row["flags"])
            if fl & HF_HAS_MIN and int(row["length_min"]) != -1 and message_length < int(row["length_min"]):
                return (0, "message is too short")
            if fl & HF_HAS_MAX and message_length > int(row["length_max"]):
                return (0, "message is too long")
        return (1, "")

    def mc2dmc(self, bit_data):
        """helpers.py:6-26"""
        if bit_data is None:
            return (-1, "no bitData provided")
        s = bit_data.replace("1", "lh").replace("0", "hl")
        return "".join("0" if s[i] == s[i + 1] else "1" for i in range(1, len(s) - 1, 2))

    def dec_2_bin_ppari(self, num):
        """helpers.py:66-88"""
        if num is None:
            return None
        nbin = format(num, "08b")
        return nbin + str(nbin.count("1") & 1)
