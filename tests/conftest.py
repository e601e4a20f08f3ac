import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def protocols():
    from pysignalduino_b200.protocol_data import load_protocol_table

    p = load_protocol_table()
    for pid, pr in p.items():
        pr.setdefault("active", True)
        pr.setdefault("name", f"Protocol_{pid}")
    return p


@pytest.fixture(scope="session")
def oracle(protocols):
    from oracle.oracle import Oracle

    return Oracle(protocols)


@pytest.fixture(scope="session")
def corpus(protocols):
    from corpus.corpus import Corpus

    return Corpus(protocols)


@pytest.fixture(scope="session")
def sdp():
    """GPU-backed drop-in class (session-wide: one engine, one table upload)."""
    from pysignalduino_b200 import SDProtocols

    return SDProtocols(device=0, mc_repaired=True)
