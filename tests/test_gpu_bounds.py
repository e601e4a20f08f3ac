"""compute-sanitizer is closed on the GPU pool, so the kernels carry their own optional bounds checks:
run every kernel from the -DSDB_BOUNDS_CHECK build (libsdb200_chk.so) and require zero violations,
with results still equal to the oracle's."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent

SCRIPT = r"""
import sys
sys.path.insert(0, %r)
from corpus.corpus import Corpus
from oracle.oracle import Oracle
from pysignalduino_b200 import SDProtocols, pack
from tests.common import compare_raw
sdp = SDProtocols(device=0, mc_repaired=True)
eng = sdp.engine()
assert eng.debug_violations(reset=True) != 0xFFFFFFFF, "not the bounds-check build"
corp, ora = Corpus(sdp.get_protocol_list()), Oracle(sdp.get_protocol_list())
for kind in (pack.KIND_MS, pack.KIND_MU):
    b = corp.pulse(kind, 30000)
    r = sdp.demodulate_packed(b)
    assert compare_raw(sdp, b, r, *ora.run_pulse_raw(b, nthreads=8)) == ""
# long messages (D up to 4096 digits: the sdb_long kernels), lines longer than the tokenizer's first pass, the format kernel
from tests.common import load_golden
recs = [r for r in load_golden("long_d.json.gz") if r["dlen"] <= pack.MAX_DIGITS]
for typ in ("MS", "MU"):
    sel = [r for r in recs if r["type"] == typ]
    st, res = sdp.demodulate_batch([r["msg"] for r in sel], typ)
    got = [(a, [[x["protocol_id"], x["payload"], x["meta"]["bit_length"]] for x in lst]) for a, lst in zip(st, res)]
    assert got == [(r["status"], r["results"]) for r in sel]
from pysignalduino_b200 import SignalParser
lines = SignalParser(protocols=sdp).parse_lines([r["line"] for r in recs[:200]])
assert sum(len(x) for x in lines) > 0
for kind in (pack.KIND_MC, pack.KIND_MN):
    b = corp.hexmsgs(kind, 30000)
    r = eng.demod_host(b, mc_repaired=True)
    assert compare_raw(sdp, b, r, *ora.run_hex_raw(b, mc_repaired=True, nthreads=8), check_bits=False) == ""
v = eng.debug_violations()
print("violations", v)
assert v == 0
"""


def test_bounds_checked_build_sees_no_violation():
    env = dict(os.environ, SDB200_CHECKED="1")
    res = subprocess.run([sys.executable, "-c", SCRIPT % str(ROOT)], env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert "violations 0" in res.stdout
