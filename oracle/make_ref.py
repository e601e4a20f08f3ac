#!/usr/bin/env python3
"""Install the UNMODIFIED reference package into the git-ignored ``oracle/_ref/`` so that it travels to the GPU box.

    python oracle/make_ref.py            (build container only: needs /root/reference; run by __graft_entry__.build())

``bench.py`` times the reference's own ``SDProtocols.demodulate`` on the GPU box's host cores (``cpu_baseline`` /
``--impl reference``) and ``oracle/ref_import.py`` falls back to this copy when ``/root/reference`` is
absent.  Nothing here is product code, and nothing under ``oracle/_ref/`` is committed (see .gitignore): the directory
is produced by ``pip install --no-deps --target oracle/_ref`` from a scratch copy of the reference tree (pip builds in
the source tree, which is read-only), i.e. the same offline install the bench contract describes for ``baseline/_ref``.
If pip cannot build the wheel, the two pure-Python packages are installed by ``shutil.copytree`` instead (same files).
"""
from __future__ import annotations

import shutil
import subprocess
import sys
import tempfile
from pathlib import Path

HERE = Path(__file__).resolve().parent
SRC = Path("/root/reference")
DST = HERE / "_ref"
PACKAGES = ("sd_protocols", "signalduino")


def installed() -> bool:
    return (DST / "sd_protocols" / "sd_protocols.py").exists() and (DST / "sd_protocols" / "protocols.json").exists()


def build(force: bool = False) -> bool:
    """Returns True when oracle/_ref holds the reference packages afterwards."""
    if installed() and not force:
        return True
    if not (SRC / "sd_protocols" / "sd_protocols.py").exists():
        return installed()
    if DST.exists():
        shutil.rmtree(DST)
    DST.mkdir(parents=True)
    how = "pip"
    with tempfile.TemporaryDirectory(prefix="sdref_") as tmp:
        work = Path(tmp) / "reference"
        shutil.copytree(SRC, work, ignore=shutil.ignore_patterns(".git", "__pycache__", "docs", "tests", "tools"))
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--quiet",
               "--find-links", "/opt/wheelhouse", "--target", str(DST), str(work)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0 or not installed():
            # the reference's pyproject.toml declares no package data, so a pip install omits sd_protocols/protocols.json
            why = (res.stderr.strip().splitlines() or ["installed without protocols.json (no package-data in pyproject.toml)"])[-1][:200]
            how = f"copytree (pip: {why})"
            shutil.rmtree(DST)
            DST.mkdir(parents=True)
            for pkg in PACKAGES:
                shutil.copytree(SRC / pkg, DST / pkg, ignore=shutil.ignore_patterns("__pycache__"))
    (DST / "HOW.txt").write_text(f"installed from {SRC} by oracle/make_ref.py via {how}\n")
    return installed()


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("oracle/_ref:", "ready" if ok else "NOT available", (DST / "HOW.txt").read_text().strip() if ok else "")
