"""Direct mcBit2* / mcRaw / mcraw calls (manchester.py:207-795, helpers.py:90-122) against reference-generated vectors.

The reference's own tests call these decoders directly (tests/test_manchester_protocols.py), so the drop-in class
exposes them; each call is one sdb_unit_mc launch on the device, the host only renders strings.
"""
import pytest

from tests.common import load_golden


def test_mc_unit_goldens_cover_every_decoder():
    groups = load_golden("mc_units.json.gz")
    seen = {}
    for g in groups:
        for m, bits, pid, mcbitnum, rc, out in g["calls"]:
            seen.setdefault(m, set()).add(rc if rc == 1 or isinstance(rc, str) else out.split(",")[0][:24])
    assert len(seen) == 13
    assert all(1 in v and len(v) >= 2 for v in seen.values()), seen
    assert "ValueError" in seen["mcBit2Funkbus"] and "TypeError" in seen["mcraw"]


@pytest.mark.gpu
def test_device_mc_units_match_reference():
    from pysignalduino_b200 import SDProtocols

    bad, total = [], 0
    for g in load_golden("mc_units.json.gz"):
        sdp = SDProtocols(device=0, mc_repaired=True)
        for pid, row in (g["patch"] or {}).items():
            sdp._protocols[pid] = dict(row)                  # the reference's tests patch the table the same way
        for m, bits, pid, mcbitnum, rc, out in g["calls"]:
            total += 1
            try:
                got = getattr(sdp, m)("golden", bits, pid, mcbitnum)
            except (TypeError, ValueError, IndexError) as e:
                got = (type(e).__name__, None)
            if tuple(got) != (rc, out):
                bad.append((m, pid, mcbitnum, len(bits), got, (rc, out)))
        sdp.engine().close()
    assert not bad, (len(bad), total, bad[:8])


def test_conv_goldens_cover_every_converter():
    seen = {}
    for m, msg, st, out in load_golden("conv_units.json.gz"):
        seen.setdefault(m, set()).add(bool(out))
    assert len(seen) == 7 and all(v == {True, False} for v in seen.values()), seen


@pytest.mark.gpu
def test_device_conv_units_match_reference(sdp):
    """Direct Conv*(msg_data) calls (helpers.py:223-716), as the reference's tests/test_helpers.py makes them."""
    bad = []
    for m, msg, st, out in load_golden("conv_units.json.gz"):
        got = getattr(sdp, m)(dict(msg))
        if got != out:
            bad.append((m, msg, got, out))
    assert not bad, (len(bad), bad[:5])
