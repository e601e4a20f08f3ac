"""Batch ``SignalParser``: firmware lines in, ``DecodedMessage`` lists out.

Mirrors ``signalduino/parser/__init__.py:17-77`` (``SignalParser.parse_line``) and adds ``parse_lines`` — the
call the GPU path exists for.  What runs where:

* framing / Mred decompression (``base.py:13-193``): host, per line (string work the firmware protocol defines);
* MS / MU (``ms.py:26-69``, ``mu.py:26-82``): the payload lines of each type go to the device in one buffer; a
  tokenizer kernel does the dict building, the MU validity regex and the demodulator's input gates
  (``csrc/sdb_lines.cu``), the demodulation kernels follow.  Lines outside the canonical grammar come back flagged
  and take the dict path (``SDProtocols.demodulate``), still decoded on the device;
* MC (``mc.py:26-93``): the reference hands ``demodulate_mc`` a dict without ``protocol_id``, so every MC line
  yields ``[]`` (``sd_protocols.py:79-83``); reproduced as is;
* MN (``mn.py:30-191``): regex, rfmode / length / regexMatch filters on the host, then all surviving
  (line x protocol) pairs in ONE device batch through the MN kernel.
"""
from __future__ import annotations

import logging
import re
from dataclasses import dataclass, field
from datetime import datetime, timezone
from typing import Any, Dict, List, Optional, Sequence

import numpy as np

from . import pack
from .capi import HIT_MM_HOST, LINE_HOSTPATH, LINE_OK, ST_OK
from .sd_protocols import SDProtocols


# ------------------------------------------------------------------------------------------ types.py:13-31
@dataclass(slots=True)
class RawFrame:
    line: str
    timestamp: datetime = field(default_factory=lambda: datetime.now(timezone.utc).replace(tzinfo=None))
    rssi: Optional[float] = None
    freq_afc: Optional[float] = None
    message_type: Optional[str] = None


@dataclass(slots=True)
class DecodedMessage:
    protocol_id: str
    payload: str
    raw: RawFrame
    metadata: dict = field(default_factory=dict)


# ------------------------------------------------------------------------------------------ base.py
def calc_rssi(raw_rssi: int) -> float:
    """base.py:207-212"""
    return ((raw_rssi - 256) / 2) - 74 if raw_rssi >= 128 else (raw_rssi / 2) - 74


def calc_afc(raw_afc: int) -> float:
    """base.py:215-220"""
    return (raw_afc - 256) / 2 if raw_afc >= 128 else raw_afc / 2


_HEX12 = re.compile(r"^[0-9A-F]{1,2}$")


def _ends_data_block(part: str) -> bool:
    """base.py:73-110: does this ';'-separated piece start a new field (and so end the reduced D= block)?"""
    c0, rest = part[0], part[1:]
    if not c0.isalpha():
        return False
    if c0 in "Dd" or ord(c0) > 127 or c0 == "M" or c0 in "om":
        return True
    if c0 in "CS" and len(rest) == 1:
        return True
    if _HEX12.match(rest.upper()):
        return True
    return c0.isalnum() and "=" in part


def decompress_payload(payload: str) -> str:
    """Mred=1 ("reduced") payload -> the plain ``key=value;`` form (base.py:13-172)."""
    if not payload.upper().startswith(("MS;", "MU;", "MO;", "MN;")):
        return payload
    if not any(ord(c) > 127 for c in payload[3:]):
        return payload
    pieces = payload.split(";")
    out: List[str] = []
    i, n = 0, len(pieces)
    while i < n:
        part = pieces[i]
        i += 1
        if not part:
            continue
        c0, rest = part[0], part[1:]
        code = ord(c0)
        if c0 in "Dd":
            # the reduced data bytes may contain ';' themselves: glue the following pieces until a field starts
            while i < n:
                nxt = pieces[i]
                if nxt and _ends_data_block(nxt):
                    break
                if nxt:
                    part += ";" + nxt
                i += 1
            digits = "".join(f"{(ord(ch) >> 4) & 0xF}{ord(ch) & 0x7}" for ch in part[1:])
            if c0 == "d":
                digits = digits[:-1]
            if digits.startswith("8"):
                digits = digits[1:]
            out.append("D=" + digits)
        elif c0 == "M":
            out.append("M" + rest.upper())
        elif code > 127:
            text = f"P{code & 7}="
            if len(rest) == 2:
                lo, hi = ord(rest[0]) & 127, ord(rest[1]) & 127
                if code & 0x20:
                    text += "-"
                if code & 0x10:
                    lo += 128
                text += str(hi * 256 + lo)
            out.append(text)
        elif c0 in "CS" and len(rest) == 1:
            out.append(f"{c0}P={rest}")
        elif c0 in "om":
            out.append(c0 + rest)
        elif rest and _HEX12.match(rest.upper()):
            out.append(f"{c0}={int(rest, 16)}")
        elif c0.isalnum():
            out.append(f"{c0}{'=' if rest else ''}{rest}")
    return ";".join(out) + ";"


def extract_payload(line: str) -> Optional[str]:
    """Payload between STX / ETX, decompressed (base.py:174-193): ``^\\x02(M[sSuUcCNOo];.*;)\\x03$``."""
    if not line:
        return None
    s = line.strip()
    if len(s) < 6 or s[0] != "\x02" or s[-1] != "\x03" or s[1] != "M" or s[2] not in "sSuUcCNOo" or s[3] != ";" \
            or s[-2] != ";" or "\n" in s:
        return None
    return decompress_payload(s[1:-1])


def _parse_to_dict(line: str) -> Dict[str, Any]:
    """ms.py:71-84 / mu.py:84-95"""
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


_MU_VALID = re.compile(r"^(?=.*D=\d+)(?:MU;(?:P[0-7]=-?[0-9]{1,5};){2,8}((?:D=\d{2,};)|(?:CP=\d;)|(?:R=\d+;)|(?:O;)|(?:e;)|(?:p;)|(?:w=\d;))*)$")
_MN_PATTERN = re.compile(r"^MN;D=(Y?)([0-9A-F]+);(?:R=([0-9]+);)?(?:A=(-?[0-9]{1,3});)?$")


class SignalParser:
    """Routes firmware lines to the decoder of their message type; same constructor as the reference."""

    def __init__(self, protocols: SDProtocols | None = None, logger: logging.Logger | None = None, rfmode: str | None = None):
        self.protocols = protocols or SDProtocols()
        self.logger = logger or logging.getLogger(__name__)
        self.protocols.register_log_callback(self._log_adapter)
        self.rfmode = rfmode
        # Lines the packed domain cannot hold (D longer than 4096 digits, more than 8 pattern slots, pulse values that are not
        # int32 ...) are NOT decoded — and never decoded differently from the reference.  Each one is logged at error level
        # and appended here as (index in the call, payload line, reason); strict_domain = True raises DomainError instead.
        self.domain_errors: List[tuple] = []
        self.strict_domain = False

    def _log_adapter(self, message: str, level: int):
        """__init__.py:54-66"""
        if level <= 1:
            self.logger.error(message)
        elif level == 2:
            self.logger.warning(message)
        elif level == 3:
            self.logger.info(message)
        else:
            self.logger.debug(message)

    # ------------------------------------------------------------------ reference API
    def parse_line(self, line: str) -> List[DecodedMessage]:
        return self.parse_lines([line])[0]

    # ------------------------------------------------------------------ batch API
    def parse_lines(self, lines: Sequence[str]) -> List[List[DecodedMessage]]:
        """``[parse_line(l) for l in lines]`` with one device batch per message type."""
        results: List[List[DecodedMessage]] = [[] for _ in lines]
        frames: List[Optional[RawFrame]] = [None] * len(lines)
        by_type: Dict[str, List[int]] = {"MS": [], "MU": [], "MC": [], "MN": []}
        now = datetime.now(timezone.utc).replace(tzinfo=None)
        for i, line in enumerate(lines):
            payload = extract_payload(line)
            if payload is None:
                continue
            mt = payload[:2].upper()
            frames[i] = RawFrame(line=payload, timestamp=now, message_type=mt)
            if mt in by_type:
                by_type[mt].append(i)
        for mt in ("MS", "MU"):
            if by_type[mt]:
                self._pulse_lines(mt, by_type[mt], frames, results)
        if by_type["MN"]:
            self._mn_lines(by_type["MN"], frames, results)
        # MC: demodulate_mc() is called without a protocol_id (mc.py:78) and returns [] (sd_protocols.py:79-83)
        return results

    def parse_lines_json(self, lines: Sequence[str]) -> List[List[str]]:
        """``[[MqttPublisher._message_to_json(m) for m in parse_line(l)] for l in lines]`` (signalduino/mqtt.py:228-245).

        MS / MU hits of lines the tokenizer takes are serialised natively in one call (``sdb_format_json``), without
        building ``DecodedMessage`` objects; everything else (host-path lines, MN) goes through ``parse_lines``."""
        out: List[List[str]] = [[] for _ in lines]
        rest: List[int] = []
        by_type: Dict[str, List[int]] = {"MS": [], "MU": []}
        payloads: List[Optional[str]] = [None] * len(lines)
        for i, line in enumerate(lines):
            payload = extract_payload(line)
            if payload is None:
                continue
            payloads[i] = payload
            mt = payload[:2].upper()
            if mt in by_type:
                by_type[mt].append(i)
            elif mt == "MN":
                rest.append(i)
        eng = self.protocols.engine()
        for mt, idx in by_type.items():
            if not idx:
                continue
            kind = pack.KIND_BY_NAME[mt]
            blobs, dev = [], []
            for i in idx:
                try:
                    blobs.append(payloads[i].encode("latin-1"))
                    dev.append(i)
                except UnicodeEncodeError:
                    rest.append(i)
            if not dev:
                continue
            lens = np.fromiter((len(b) for b in blobs), dtype=np.int64, count=len(blobs))
            offs = np.zeros(len(blobs), dtype=np.int64)
            np.cumsum(lens[:-1] + 1, out=offs[1:])
            text = np.frombuffer(b"\n".join(blobs) + b"\n", dtype=np.uint8)
            offs32 = offs.astype(np.uint32)
            res, info = eng.demod_lines(kind, text, offs32, lens.astype(np.uint32))
            rest.extend(dev[int(k)] for k in np.nonzero(info["status"] == LINE_HOSTPATH)[0])
            if len(res.hits):
                pool, off = eng.format_json(kind, res.hits, res.bits, text, offs32, info)
                js = pool.decode("ascii")
                o = res.out
                take = (o["nhits"] > 0) & (o["status"] == ST_OK) & (info["status"] == LINE_OK)
                mmh = (res.hits["flags"] & HIT_MM_HOST) != 0
                if mmh.any():                                               # host-evaluated modulematch: those lines take parse_lines
                    lines_mm = np.unique(res.hits["msg"][mmh].astype(np.int64))
                    take[lines_mm] = False
                    rest.extend(dev[int(k)] for k in lines_mm)
                for k in np.nonzero(take)[0]:
                    h0, nh = int(o["hit_off"][k]), int(o["nhits"][k])
                    out[dev[int(k)]] = [js[int(off[h]) : int(off[h + 1])] for h in range(h0, h0 + nh)]
        if rest:
            import json

            rest.sort()
            for i, msgs in zip(rest, self.parse_lines([lines[i] for i in rest])):
                out[i] = [json.dumps({"protocol_id": m.protocol_id, "payload": m.payload, "metadata": m.metadata}, indent=4) for m in msgs]
        return out

    def parse_text_json(self, raw: bytes):
        """A raw receive buffer ('\\n'-separated firmware lines, bytes) -> the MQTT JSON of every decoded message, without
        any per-line Python work on the volume path: native framing / decompression (``sdb_frame_lines_inplace``), tokenizer +
        demodulation kernels (``sdb_demod_lines_host``), native JSON (``sdb_format_json``).

        Returns ``(batches, extra)``: ``batches`` = one ``(json_pool: bytes, str_off: uint64[nhits + 1], hit_line:
        int64[nhits])`` per message type with device hits (hit k is ``json_pool[str_off[k]:str_off[k + 1]]`` and belongs to
        raw line ``hit_line[k]``; the hits of one line are consecutive and in reference order); ``extra`` = ``{raw line
        index: [json, ...]}`` for the lines that took the Python path (MN, lines outside the tokenizer's grammar)."""
        from .capi import FRAME_NONE, FRAME_PYPATH, FRAME_SIDE, frame_chunks

        rawbuf = np.frombuffer(raw, dtype=np.uint8)
        eng = self.protocols.engine()
        batches = []
        slow: List[int] = []
        for byte_base, byte_end, line_base, off, ln, typ, side in frame_chunks(raw):       # chunk k + 1 is framed while chunk k decodes
            framed = typ != FRAME_NONE
            slow.extend(line_base + int(i) for i in np.nonzero(framed & (((typ & FRAME_PYPATH) != 0) | ((typ & 0x0F) == pack.KIND_MN)))[0])
            for kind in (pack.KIND_MS, pack.KIND_MU):
                # plain payloads are addressed inside the caller's buffer, decompressed ones inside the side buffer
                for text, sel in ((rawbuf[byte_base:byte_end], np.nonzero(typ == kind)[0]), (side, np.nonzero(typ == (kind | FRAME_SIDE))[0])):
                    if not len(sel):
                        continue
                    loff = np.ascontiguousarray(off[sel])
                    res, info = eng.demod_lines(kind, text, loff, np.ascontiguousarray(ln[sel]))
                    slow.extend(line_base + int(i) for i in sel[info["status"] == LINE_HOSTPATH])
                    if len(res.hits):
                        hits = res.hits
                        mmh = (hits["flags"] & HIT_MM_HOST) != 0
                        if mmh.any():                                       # host-evaluated modulematch: those lines take parse_lines
                            lines_mm = np.unique(hits["msg"][mmh].astype(np.int64))
                            slow.extend(line_base + int(sel[k]) for k in lines_mm)
                            hits = hits[~np.isin(hits["msg"].astype(np.int64), lines_mm)]
                        if len(hits):
                            pool, soff = eng.format_json(kind, hits, res.bits, text, loff, info)
                            batches.append((pool, soff, line_base + sel[hits["msg"].astype(np.int64)]))
        extra: Dict[int, List[str]] = {}
        if slow:
            import json

            slow.sort()
            rows = raw.split(b"\n")
            for i, msgs in zip(slow, self.parse_lines([rows[i].decode("latin-1") for i in slow])):
                if msgs:
                    extra[i] = [json.dumps({"protocol_id": m.protocol_id, "payload": m.payload, "metadata": m.metadata}, indent=4)
                                for m in msgs]
        return batches, extra

    # ------------------------------------------------------------------ MS / MU
    def _pulse_lines(self, mt: str, idx: List[int], frames, results) -> None:
        kind = pack.KIND_BY_NAME[mt]
        eng = self.protocols.engine()
        dev: List[int] = []                 # positions (in idx) the tokenizer takes
        blobs: List[bytes] = []
        slow: List[int] = []
        for j, i in enumerate(idx):
            try:
                blobs.append(frames[i].line.encode("latin-1"))
                dev.append(j)
            except UnicodeEncodeError:
                slow.append(i)
        if dev:
            lens = np.fromiter((len(b) for b in blobs), dtype=np.int64, count=len(blobs))
            offs = np.zeros(len(blobs), dtype=np.int64)
            np.cumsum(lens[:-1] + 1, out=offs[1:])
            text = np.frombuffer(b"\n".join(blobs) + b"\n", dtype=np.uint8)
            res, info = eng.demod_lines(kind, text, offs.astype(np.uint32), lens.astype(np.uint32))
            hostpath = np.nonzero(info["status"] == LINE_HOSTPATH)[0]
            slow.extend(idx[dev[int(k)]] for k in hostpath)
            self._emit_pulse(kind, [idx[j] for j in dev], blobs, res, info, frames, results)
        for i in sorted(slow):
            self._dict_path(mt, i, frames, results)

    def _emit_pulse(self, kind: int, line_idx: List[int], blobs: List[bytes], res, info, frames, results) -> None:
        eng = self.protocols.engine()
        ids = eng.table.ids
        if not len(res.hits):
            return
        pool, off = eng.format_hits(kind, res.hits, res.bits)
        text = pool.decode("latin-1")
        out, hits = res.out, res.hits
        clk_cache: Dict[int, float] = {}
        for k in np.nonzero((out["nhits"] > 0) & (out["status"] == ST_OK) & (info["status"] == LINE_OK))[0]:
            k = int(k)
            i = line_idx[k]
            frame = frames[i]
            li = info[k]
            rssi_txt = None
            if li["has_r"]:
                a = int(li["r_off"])
                rssi_txt = blobs[k][a : a + int(li["r_len"])].decode("latin-1")
                frame.rssi = calc_rssi(int(rssi_txt))                       # ms.py:88-92 (R passed the digit gate)
            if kind == pack.KIND_MS and b"F" in blobs[k]:
                d = _parse_to_dict(frame.line)                              # rare: AFC field (ms.py:94-98)
                if "F" in d:
                    try:
                        frame.freq_afc = calc_afc(int(d["F"]))
                    except (ValueError, TypeError):
                        self.logger.warning("Could not parse AFC value: %s", d["F"])
            h0 = int(out["hit_off"][k])
            lst = results[i]
            for hi in range(h0, h0 + int(out["nhits"][k])):
                h = hits[hi]
                pi = int(h["proto"])
                if int(h["flags"]) & HIT_MM_HOST:                           # user-edited table: regex the device program cannot express
                    if not re.search(str(self.protocols.get_property(ids[pi], "modulematch")), text[int(off[hi]) : int(off[hi + 1])]):
                        continue
                if kind == pack.KIND_MS:
                    clock = float(li["clock"])                              # message_synced.py:239
                else:
                    if pi not in clk_cache:
                        clk_cache[pi] = float(self.protocols.check_property(ids[pi], "clockabs", 1))
                    clock = clk_cache[pi]                                   # message_unsynced.py:288
                lst.append(DecodedMessage(protocol_id=str(ids[pi]), payload=text[int(off[hi]) : int(off[hi + 1])], raw=frame,
                                          metadata={"bit_length": int(h["nbits"]), "rssi": rssi_txt, "clock": clock}))

    def _dict_path(self, mt: str, i: int, frames, results) -> None:
        """One line through the dict API (ms.py:38-69 / mu.py:48-82): anything the tokenizer does not take."""
        frame = frames[i]
        if mt == "MU" and not _MU_VALID.match(frame.line):
            return
        msg = _parse_to_dict(frame.line)
        if "D" not in msg:
            return
        msg["data"] = msg["D"]
        for key, fn, attr in (("R", calc_rssi, "rssi"), ("F", calc_afc, "freq_afc")):
            if key in msg:
                try:
                    setattr(frame, attr, fn(int(msg[key])))
                except (ValueError, TypeError):
                    self.logger.warning("Could not parse %s value: %s", key, msg[key])
        try:
            decoded = self.protocols.demodulate(msg, mt)
        except pack.DomainError as e:
            # NOT a reference outcome: the packed domain cannot hold this line, so it was not decoded.  Reported apart from
            # the reference's own "logged and dropped" exceptions so that a caller can see (and count) every such line.
            self._domain_error(i, frame.line, str(e))
            return
        except Exception:                                                   # ms.py:52-54: logged and dropped
            self.logger.exception("Error during %s demodulation for line: %s", mt, frame.line)
            return
        for d in decoded:
            results[i].append(DecodedMessage(protocol_id=str(d["protocol_id"]), payload=str(d.get("payload", "")), raw=frame,
                                             metadata=d.get("meta", {})))

    def _domain_error(self, i: int, line: str, why: str) -> None:
        self.domain_errors.append((i, line, why))
        self.logger.error("line outside the packed domain, NOT decoded (%s): %s", why, line[:200])
        if self.strict_domain:
            raise pack.DomainError(f"line {i}: {why}")

    # ------------------------------------------------------------------ MN (mn.py:30-191)
    def _mn_lines(self, idx: List[int], frames, results) -> None:
        P = self.protocols
        mn_ids = P.get_keys("modulation")
        tasks = []                      # (line index, pid, raw_data, rssi, freq_afc, modulation, proto_rfmode, has_method)
        msgs = []
        for i in idx:
            m = _MN_PATTERN.match(frames[i].line)
            if not m:
                continue
            raw = m.group(2)
            rssi = calc_rssi(int(m.group(3))) if m.group(3) else None
            afc = round((26000000 / 16384 * int(m.group(4)) / 1000), 0) if m.group(4) else None
            for pid in mn_ids:
                prf = P.check_property(pid, "rfmode", None)
                if not prf or (self.rfmode and prf != self.rfmode):
                    continue
                if not P.length_in_range(pid, len(raw))[0]:
                    continue
                rx = P.check_property(pid, "regexMatch", None)
                if rx and not re.search(rx, raw):
                    continue
                method = P.get_property(pid, "method")
                if method:
                    name = method.split(".")[-1]
                    fn = getattr(P, name, None)
                    if not (fn and callable(fn)):
                        self.logger.warning("MN Parse: Method %s not found for protocol %s", name, pid)
                        continue
                    msgs.append({"data": raw, "protocol_id": pid})
                tasks.append((i, pid, raw, rssi, afc, P.check_property(pid, "modulation", None), prf, bool(method)))
        statuses, decoded = P.demodulate_batch(msgs, "MN") if msgs else ([], [])
        k = 0
        for i, pid, raw, rssi, afc, modulation, prf, has_method in tasks:
            payload = raw
            if has_method:
                st, lst = statuses[k], decoded[k]
                k += 1
                if st == "DomainError":                                     # e.g. D longer than 512 hex characters: not decoded
                    self._domain_error(i, frames[i].line, f"MN protocol {pid}: D outside the packed domain")
                    continue
                if st != "ok":                                              # mn.py:167-169: logged, next protocol
                    self.logger.error("Error executing method for protocol %s: %s", pid, st)
                    continue
                # mn.py:150-166: a non-empty list of dicts gives its payload, anything else str(result) -> "[]"
                payload = lst[0].get("payload", raw) if lst else "[]"
            results[i].append(DecodedMessage(protocol_id=str(pid), payload=f"{P.check_property(pid, 'preamble', '')}{payload}",
                                             raw=frames[i], metadata={"rssi": rssi, "freq_afc": afc, "modulation": modulation,
                                                                      "rfmode": prf}))
