"""Time the REAL reference — ``SDProtocols.demodulate`` of PySignalduino, unmodified — on the host cores.

BASELINE.md §4: the reference's own Python path in ``multiprocessing.Pool(N)`` with one ``SDProtocols()`` per worker,
N = ``os.cpu_count()``, on exactly the dict inputs that are packed for the GPU (dict construction excluded on both sides),
over a seed-stratified sample of the benchmark corpus.  The reference comes from ``/root/reference`` in the build container
and from the copy ``oracle/make_ref.py`` installs into the git-ignored ``oracle/_ref/`` everywhere else (GPU box).

Test / measurement infrastructure: only bench.py's ``cpu_baseline`` and ``--impl reference`` legs use it.  MC runs in the
"repaired" mode of SURVEY §8c (the two documented one-line fixes; as shipped every MC call raises TypeError), which is what
the GPU path is benchmarked in.
"""
from __future__ import annotations

import os
import sys
import time
from pathlib import Path
from typing import Any, Dict, List, Sequence, Tuple

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

_REF = None


def available() -> bool:
    from oracle import ref_import

    return ref_import.available()


def _init_worker() -> None:
    global _REF
    import logging

    logging.disable(logging.CRITICAL)
    from oracle import ref_import

    _REF = ref_import.repaired_class()()


def _decode_chunk(task: Tuple[str, List[Dict[str, Any]]]):
    """One worker call: demodulate a list of dicts of one message type; canonical results for the parity check."""
    typ, msgs = task
    out = []
    ref = _REF
    for m in msgs:
        try:
            res = ref.demodulate(m, typ)
            out.append(("ok", [(str(r["protocol_id"]), str(r["payload"]), int((r.get("meta") or {}).get("bit_length", -1))) for r in res]))
        except Exception as e:  # noqa: BLE001 - the exception type is the reference's outcome
            out.append((type(e).__name__, []))
    return out


class ReferencePool:
    """``Pool(workers)`` of reference instances (spawned: no CUDA state is inherited)."""

    def __init__(self, workers: int | None = None):
        import multiprocessing as mp

        self.workers = workers or os.cpu_count() or 1
        self.pool = mp.get_context("spawn").Pool(self.workers, initializer=_init_worker)
        self.pool.map(_noop, range(self.workers * 2))            # workers up, reference imported, table loaded

    def decode(self, typ: str, msgs: Sequence[Dict[str, Any]], chunk: int = 128):
        """-> (canonical results in order, wall seconds of the decode alone)."""
        tasks = [(typ, list(msgs[i : i + chunk])) for i in range(0, len(msgs), chunk)]
        t0 = time.perf_counter()
        parts = self.pool.map(_decode_chunk, tasks, chunksize=1)
        dt = time.perf_counter() - t0
        return [r for p in parts for r in p], dt

    def close(self) -> None:
        self.pool.close()
        self.pool.join()


def _noop(_):
    return _REF is not None
