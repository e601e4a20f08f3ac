#!/bin/bash
# everything that travels to the GPU box as a built artefact: libsdb200.so, libsdb200_chk.so, _fastpack.so, the oracle, the corpus
# generator, oracle/_ref.  Run before every gpurun call (stale .so files are what the box would otherwise execute).
cd "$(dirname "$0")/.." && python __graft_entry__.py
