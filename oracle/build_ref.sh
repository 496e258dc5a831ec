#!/bin/bash
# Compiles the few reference sources that build standalone FROM
# WHERE THEY LIE under /root/reference into oracle/_ref/ (git-ignored, travels
# with gpurun).  No reference source is copied into the repo.  The full
# reference (src/Makefile*) needs MPI + Eigen + Boost, none installed: unbuildable.
set -e
REF=${1:-/root/reference}
OUT="$(cd "$(dirname "$0")" && pwd)/_ref"
mkdir -p "$OUT"
CXX=/usr/bin/g++
# LUT generator: its stdout is the text of src/dotp_lut.h (tests/golden/make_golden.py runs it and compares; nothing is stored)
$CXX -O1 -w -o "$OUT/mk_lut" "$REF/src/mk_lut.cpp"
rm -f "$OUT/dotp_lut_generated.h"
# ARMS sampler (BayesW): libc only. rand() is redirected to the shim at link time (--wrap=rand) so that tests control the uniforms.
HERE="$(cd "$(dirname "$0")" && pwd)"
$CXX -O2 -w -fPIC -shared -fno-fast-math -ffp-contract=off -Wl,--wrap=rand -o "$OUT/libarms_ref.so" "$REF/src/BayesW_arms.cpp" "$HERE/arms_shim.cpp"
echo "built: $(ls "$OUT")"
