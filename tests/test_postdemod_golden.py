"""postDemo_* against reference-generated vectors: the C oracle on CPU, the device functions on GPU."""
import ctypes as C

import numpy as np
import pytest

from tests.common import load_golden

PD = {"postDemo_EM": 1, "postDemo_Revolt": 2, "postDemo_FS20": 3, "postDemo_FHT80": 4, "postDemo_FHT80TF": 5,
      "postDemo_WS2000": 6, "postDemo_WS7035": 7, "postDemo_WS7053": 8, "postDemo_lengtnPrefix": 9}


def test_oracle_postdemod_matches_reference():
    from oracle.oracle import lib

    L = lib()
    L.ora_postdemod.restype = C.c_int
    bad = 0
    for r in load_golden("postdemod.json.gz"):
        a = np.asarray(r["bits"], dtype=np.uint8)
        out = np.zeros(len(a) + 64, dtype=np.uint8)
        n_out = C.c_int(0)
        rc = L.ora_postdemod(PD[r["method"]], C.c_void_p(a.ctypes.data) if len(a) else None, len(a),
                             C.c_void_p(out.ctypes.data), len(out), C.byref(n_out))
        if isinstance(r["rc"], str):
            ok = rc == -2 and r["rc"] == "ValueError"
        elif r["rc"] == 1:
            ok = rc == 1 and out[: n_out.value].tolist() == r["out"]
        else:
            ok = rc == 0
        bad += not ok
    assert bad == 0


@pytest.mark.gpu
def test_device_postdemod_matches_reference(sdp):
    """The drop-in class' postDemo_* methods run the device functions (sdb_unit_postdemod)."""
    bad = []
    for i, r in enumerate(load_golden("postdemod.json.gz")):
        meth = getattr(sdp, r["method"])
        if isinstance(r["rc"], str):
            with pytest.raises(ValueError):
                meth("t", r["bits"])
            continue
        rc, out = meth("t", r["bits"])
        if r["rc"] == 1:
            ok = rc == 1 and out == r["out"]
        else:
            ok = rc == 0 and out is None
        if not ok:
            bad.append((i, r["method"], len(r["bits"]), rc))
    assert not bad, bad[:10]
