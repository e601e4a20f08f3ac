"""Build libsdb200.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension machinery).

    python -m pysignalduino_b200.build_ext [--force] [--verbose]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libsdb200.so"
LIB_CHK = PKG / "libsdb200_chk.so"      # same sources with -DSDB_BOUNDS_CHECK
HEADERS_EXTRA = ["sdb_pulse.cu"]
SOURCES = ["sdb_capi.cu", "sdb_pulse.cu", "sdb_pulse_long.cu", "sdb_hex.cu", "sdb_lines.cu", "sdb_frame.cu", "sdb_format.cu"]
HEADERS = ["sdb_table.h", "sdb_pulse.h", "sdb_fmt.h", "sdb_postdemod.cuh", "sdb_pyctype.h", "../../include/sdb200.h"]
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]


def nvcc_path() -> str:
    cand = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(cand).exists():
        raise RuntimeError("nvcc not found: libsdb200.so cannot be built (there is no CPU fallback)")
    return cand


def needs_build(lib: Path = LIB) -> bool:
    if not lib.exists():
        return True
    t = lib.stat().st_mtime
    return any((CSRC / f).stat().st_mtime > t for f in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False, check: bool = False, defines=(), out: Path | None = None) -> Path:
    """Build the library; ``defines`` / ``out`` produce experiment variants (``-DNAME[=value]`` into another file, selected at
    run time with SDB200_LIB=<path>)."""
    lib = out if out is not None else (LIB_CHK if check else LIB)
    if not force and not needs_build(lib):
        return lib
    cmd = [nvcc_path(), "-O3", "-std=c++17", "-lineinfo", *ARCH_FLAGS, "-Xcompiler", "-fPIC", "-shared",
           "-Xptxas", "-v" if verbose else "-O3", "-o", str(lib)] + (["-DSDB_BOUNDS_CHECK"] if check else []) \
        + [f"-D{d}" for d in defines] + [str(CSRC / s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed building {lib.name}")
    return lib


FASTPACK = PKG / "_fastpack.so"


def build_fastpack(force: bool = False) -> Path:
    """The native dict packer (CPython extension, host code only): gcc, in-tree, next to libsdb200.so."""
    import sysconfig

    src = CSRC / "sdb_fastpack.c"
    if not force and FASTPACK.exists() and FASTPACK.stat().st_mtime >= src.stat().st_mtime:
        return FASTPACK
    cc = shutil.which("gcc") or shutil.which("cc")
    if not cc:
        raise RuntimeError("gcc not found: _fastpack.so cannot be built")
    cmd = [cc, "-O2", "-shared", "-fPIC", "-Wall", f"-I{sysconfig.get_paths()['include']}", "-o", str(FASTPACK), str(src)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("gcc failed building _fastpack.so")
    return FASTPACK


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a[6:] for a in sys.argv[1:] if a.startswith("--out=")]
    if not outs:
        print(build_fastpack(force="--force" in sys.argv))
    print(build(force="--force" in sys.argv or bool(defs), verbose="--verbose" in sys.argv, check="--check" in sys.argv, defines=defs,
                out=Path(outs[0]).resolve() if outs else None))
