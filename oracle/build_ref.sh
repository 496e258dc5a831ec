#!/bin/bash
# Compiles the few reference sources that build standalone FROM
# WHERE THEY LIE under /root/reference into oracle/_ref/ (git-ignored, travels
# with gpurun).  No reference source is copied into the repo.  The full
# reference (src/Makefile*) needs MPI + Eigen + Boost, none installed: unbuildable.
set -e
REF=${1:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
mkdir -p "$OUT"
CXX=/usr/bin/g++
# LUT generator: its stdout is the text of src/dotp_lut.h (tests/golden/make_golden.py runs it and compares; nothing is stored)
$CXX -O1 -w -o "$OUT/mk_lut" "$REF/src/mk_lut.cpp"
rm -f "$HERE/_ref/dotp_lut_generated.h"
# ARMS sampler (BayesW): libc only. rand() is redirected to the shim at link time (--wrap=rand) so that tests control the uniforms.
$CXX -O2 -w -fPIC -shared -fno-fast-math -ffp-contract=off -Wl,--wrap=rand -o "$OUT/libarms_ref.so" "$REF/src/BayesW_arms.cpp" "$HERE/arms_shim.cpp"

# Function bodies of the reference's unit kernels and staging code, cut out by line range into oracle/_ref/*.inc (git-ignored)
# and compiled behind the stub declarations of oracle/ref_kernels_shim.cpp (pins sparse_dotprod, sparse_scaadd,
# the two LUT statements of the marker loop, sparse_data_fill_indices, the NA compaction and get_bed_marker_from_sparse to the
# reference's own object code). The md5 sums guard the line numbers.
if echo "7b13ffb04c5b09169643d2a7c95eb7d2  $REF/src/BayesRRm.cpp" | md5sum -c --status && echo "15253b99ee58214a6f000ed1e622efd7  $REF/src/data.cpp" | md5sum -c --status; then
  sed -n '60,388p'    "$REF/src/BayesRRm.cpp" > "$HERE/_ref/brr_kernels.inc"
  sed -n '396,413p'   "$REF/src/BayesRRm.cpp" > "$HERE/_ref/brr_blocks.inc"
  sed -n '1757,1849p' "$REF/src/BayesRRm.cpp" > "$HERE/_ref/brr_num.inc"
  sed -n '1976,2019p' "$REF/src/BayesRRm.cpp" > "$HERE/_ref/brr_deps.inc"
  sed -n '826,865p'   "$REF/src/data.cpp"     > "$HERE/_ref/data_bed.inc"
  sed -n '1112,1290p' "$REF/src/data.cpp"     > "$HERE/_ref/data_sparse.inc"
  # strict IEEE, no OpenMP: the bodies' sums then run in source order (the oracle's strict build does the same)
  $CXX -std=gnu++17 -O2 -w -mavx2 -fPIC -shared -fvisibility=hidden -fno-fast-math -ffp-contract=off -I "$REF/src" -I "$HERE" \
       -o "$OUT/libref_kernels.so" "$HERE/ref_kernels_shim.cpp"
  for f in brr_kernels brr_blocks brr_num brr_deps data_bed data_sparse; do rm -f "$HERE/_ref/$f.inc"; done   # cut-outs are build intermediates only
else
  echo "reference sources differ from the surveyed checkout: libref_kernels.so not built" >&2
fi
# format oracles: the reference's own readers of .bet / .cpn / .eps (postproc/, libc only)
for t in beta_converter components_converter epsilon_converter; do
  $CXX -O1 -w -o "$OUT/$t" "$REF/postproc/$t.cpp"
done
echo "built: $(ls "$OUT")"
