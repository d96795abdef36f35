#!/usr/bin/env bash
# TEST INFRASTRUCTURE. Builds the reference's own CUDA sources (unmodified, read from
# /root/reference where they lie) plus our C-ABI shims into oracle/_ref/*.so.
# Flags mirror what torch's BuildExtension passes for the reference's setup.py
# (no fast-math; nvcc defaults -fmad=true -prec-div=true -prec-sqrt=true), plus the two
# force-included headers gcc 13 needs. oracle/_ref/ is git-ignored but
# travels to the GPU box.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${REF_ROOT:-/root/reference}"
RAST="$REF/submodules/depth-diff-gaussian-rasterization"
KNN="$REF/submodules/simple-knn"
OUT="$HERE/_ref"
mkdir -p "$OUT"
if [ ! -d "$RAST/cuda_rasterizer" ]; then
  echo "reference sources not present at $REF; keeping prebuilt oracle/_ref" >&2
  exit 0
fi
COMMON="-O3 -std=c++17 --expt-relaxed-constexpr -gencode arch=compute_100,code=sm_100 -include cstdint -include cfloat -Xcompiler -fPIC -shared -w"
build_rast() {
  nvcc $COMMON -I"$RAST" -I"$RAST/third_party/glm" -o "$OUT/libref_rast.so" \
    "$RAST/cuda_rasterizer/rasterizer_impl.cu" "$RAST/cuda_rasterizer/forward.cu" \
    "$RAST/cuda_rasterizer/backward.cu" "$HERE/ref_shim_rast.cu"
}
build_knn() {
  nvcc $COMMON -I"$KNN" -o "$OUT/libref_knn.so" "$KNN/simple_knn.cu" "$HERE/ref_shim_knn.cu"
}
[ "$OUT/libref_rast.so" -nt "$HERE/ref_shim_rast.cu" ] || build_rast &
[ "$OUT/libref_knn.so" -nt "$HERE/ref_shim_knn.cu" ] || build_knn &
wait
ls -la "$OUT"
