"""Long D streams (the reference has no length limit, message_unsynced.py:22-25): reference-generated goldens with D of
1023 .. 5000 digits.  Up to SDB_MAX_DIGITS = 4096 the result must equal the reference's; beyond that the message — and
only that message — reports DomainError (it is not decoded, never decoded differently)."""
import json

import numpy as np
import pytest

from pysignalduino_b200 import pack
from tests.common import diff_report, golden_expected, load_golden


def _expected(r):
    if r["dlen"] > pack.MAX_DIGITS:
        return ("DomainError", [])
    return golden_expected(r)


def test_oracle_matches_reference_long(oracle):
    """CPU: the C oracle on every in-domain record (it is the checker of the full-size GPU runs)."""
    recs = [r for r in load_golden("long_d.json.gz") if r["dlen"] <= pack.MAX_DIGITS]
    assert {r["dlen"] for r in recs} >= {1023, 1024, 1025, 1500, 2000, 2048, 4096}
    for typ in ("MS", "MU"):
        sel = [r for r in recs if r["type"] == typ]
        got = oracle.run_pulse(pack.pack_pulse([r["msg"] for r in sel], pack.KIND_BY_NAME[typ], strict=True))
        exp = [golden_expected(r) for r in sel]
        assert got == exp, diff_report(got, exp)
        assert sum(len(e[1]) for e in exp) > 100


@pytest.mark.gpu
@pytest.mark.parametrize("typ", ["MS", "MU"])
def test_gpu_demodulate_batch_long(sdp, typ):
    """demodulate_batch over all lengths at once: fast kernels, long kernels and out-of-domain records in ONE batch."""
    recs = [r for r in load_golden("long_d.json.gz") if r["type"] == typ]
    short = [r for r in load_golden("corpus_ms.json.gz" if typ == "MS" else "corpus_mu.json.gz")[:300]]
    allr = []
    for i, r in enumerate(recs):                  # interleave with ordinary messages: both kernel families see a mixed launch
        allr.append(r)
        allr.append(dict(short[i % len(short)], dlen=0))
    st, res = sdp.demodulate_batch([r["msg"] for r in allr], typ)
    got = [(s, [(str(x["protocol_id"]), x["payload"], int(x["meta"]["bit_length"])) for x in lst]) for s, lst in zip(st, res)]
    exp = [_expected(r) for r in allr]
    assert got == exp, diff_report(got, exp)
    assert sum(1 for s in st if s == "DomainError") == sum(1 for r in recs if r["dlen"] > pack.MAX_DIGITS) > 0
    assert sum(len(e[1]) for e, r in zip(exp, allr) if r["dlen"] > 1024) > 50      # long-kernel accepts, not only rejects


@pytest.mark.gpu
def test_gpu_scalar_long_and_domain(sdp):
    from pysignalduino_b200 import DomainError

    recs = load_golden("long_d.json.gz")
    for r in [x for x in recs if x["dlen"] in (1025, 2000, 4096)][:12]:
        if r["status"] != "ok":
            with pytest.raises(IndexError):
                sdp.demodulate(r["msg"], r["type"])
            continue
        out = sdp.demodulate(r["msg"], r["type"])
        assert [[o["protocol_id"], o["payload"], o["meta"]["bit_length"]] for o in out] == r["results"]
    big = next(x for x in recs if x["dlen"] == 4097)
    with pytest.raises(DomainError):
        sdp.demodulate(big["msg"], big["type"])


@pytest.mark.gpu
def test_gpu_parse_lines_long(sdp):
    """SignalParser.parse_lines on the firmware lines (tokenizer: first pass up to 1280 bytes, second pass up to 4608)."""
    from pysignalduino_b200 import SignalParser

    recs = load_golden("long_d.json.gz")
    p = SignalParser(protocols=sdp)
    got = p.parse_lines([r["line"] for r in recs])
    bad = []
    for i, (r, g) in enumerate(zip(recs, got)):
        exp = r["line_results"] if r["dlen"] <= pack.MAX_DIGITS else []
        gg = [{"protocol_id": m.protocol_id, "payload": m.payload, "metadata": m.metadata} for m in g]
        if gg != exp:
            bad.append((i, r["type"], r["dlen"], len(gg), len(exp)))
    assert not bad, bad[:10]
    n_out = sum(1 for r in recs if r["dlen"] > pack.MAX_DIGITS)
    assert len(p.domain_errors) == n_out and n_out > 0          # reported one by one, not dropped silently
    p.strict_domain = True
    with pytest.raises(pack.DomainError):
        p.parse_lines([next(r["line"] for r in recs if r["dlen"] > pack.MAX_DIGITS)])


@pytest.mark.gpu
def test_gpu_parse_text_json_long(sdp):
    """Raw receive buffer -> MQTT JSON (native framing, tokenizer, demodulation, JSON) on the same lines."""
    from pysignalduino_b200 import SignalParser

    recs = [r for r in load_golden("long_d.json.gz") if r["dlen"] <= pack.MAX_DIGITS]
    raw = ("\n".join(r["line"] for r in recs) + "\n").encode("latin-1")
    p = SignalParser(protocols=sdp)
    batches, extra = p.parse_text_json(raw)
    per_line = {i: list(v) for i, v in extra.items()}
    for pool, soff, hit_line in batches:
        for k in range(len(hit_line)):
            per_line.setdefault(int(hit_line[k]), []).append(pool[int(soff[k]) : int(soff[k + 1])].decode("ascii"))
    bad = []
    for i, r in enumerate(recs):
        exp = [json.dumps({"protocol_id": x["protocol_id"], "payload": x["payload"], "metadata": x["metadata"]}, indent=4) for x in r["line_results"]]
        if per_line.get(i, []) != exp:
            bad.append((i, r["type"], r["dlen"], len(per_line.get(i, [])), len(exp)))
    assert not bad, bad[:10]


@pytest.mark.gpu
def test_gpu_one_bad_message_does_not_abort_the_batch(sdp, corpus, oracle):
    """A batch with unrepresentable messages still returns every other message (bit-exact against the oracle)."""
    from tests.common import canonical_gpu

    batch = corpus.pulse(pack.KIND_MU, 4000)
    msgs = [pack.unpack_pulse(batch, i) for i in range(batch.n)]
    bad_at = [7, 1999, 3999]
    msgs[7] = dict(msgs[7], data="0" * 5000)
    msgs[1999] = dict(msgs[1999], P3="123.5")
    msgs[3999] = {f"P{k}": "100" for k in range(10)} | {"data": "0101"}
    st, res = sdp.demodulate_batch(msgs, "MU")
    exp = oracle.run_pulse(batch)
    for i in range(batch.n):
        if i in bad_at:
            assert st[i] == "DomainError" and res[i] == []
        else:
            got = (st[i], [(str(x["protocol_id"]), x["payload"], int(x["meta"]["bit_length"])) for x in res[i]])
            assert got == exp[i], (i, got, exp[i])
