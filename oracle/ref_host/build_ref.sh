#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY (oracle/).
# Host-compile the real reference (read-only mount /root/reference) into
# oracle/_ref/libsrt_ref.so.  The reference's own build system (CMake + nvcc) is NOT used:
# its hot path compiles from its own few sources with g++ once CUDA's qualifiers and runtime
# are shimmed (shim/cuda_shim.h).  Patched scratch copies live in a temp dir outside the repo
# and are removed afterwards; only the .so is kept (git-ignored, travels with gpurun).
#
#   oracle/ref_host/build_ref.sh [--draw-order ltr|rtl] [--physical]
# --physical additionally patches materials/material.cuh:67 (sellmeier_C[i] = c[i]) and names the
# library libsrt_ref_<order>_physical.so: the reference with physically meant glass.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${SRT_REFERENCE_DIR:-/root/reference}"
OUT="$HERE/../_ref"
ORDER="ltr"
PHYS=""
SUFFIX=""
while [[ $# -gt 0 ]]; do
  case "$1" in
    --draw-order) ORDER="$2"; shift 2 ;;
    --physical) PHYS="--physical"; SUFFIX="_physical"; shift ;;
    *) echo "build_ref.sh: unknown argument $1" >&2; exit 2 ;;
  esac
done
if [[ ! -d "$REF" ]]; then
  echo "build_ref.sh: $REF not present (GPU box?) -- keeping prebuilt oracle/_ref" >&2
  exit 0
fi
mkdir -p "$OUT"
STAGE="$(mktemp -d /tmp/srt_ref_stage.XXXXXX)"
trap 'rm -rf "$STAGE"' EXIT
python3 "$HERE/patch_ref.py" --ref "$REF" --out "$STAGE/src" --draw-order "$ORDER" $PHYS

INC=("-I$STAGE/src")
for d in materials primitives bvh utils rendering refraction color spectrum ray math io scene _log_; do
  INC+=("-I$STAGE/src/$d")
done
CXXFLAGS=(-std=c++20 -O2 -ffp-contract=off -fPIC -fopenmp -w -x c++ -include "$HERE/shim/cuda_shim.h" -I"$HERE/shim" "${INC[@]}")
SRCS=(utils/cie_const.cu utils/color_const.cu utils/cuda_utility.cu spectrum/spectrum.cu color/color.cu
      refraction/sellmeier.cu primitives/transform.cu primitives/tri.cu primitives/tri_quad.cu
      primitives/tri_box.cu primitives/prism.cu primitives/pyramid.cu bvh/aabb.cu bvh/bvh.cu
      materials/material.cu rendering/camera.cu rendering/rendering.cu rendering/render_manager.cu
      scene/scene.cu io/params.cpp _log_/log_context.cpp)
OBJS=()
pids=()
for s in "${SRCS[@]}"; do
  o="$STAGE/$(echo "$s" | tr '/.' '__').o"
  OBJS+=("$o")
  g++ "${CXXFLAGS[@]}" -c "$STAGE/src/$s" -o "$o" &
  pids+=($!)
done
g++ "${CXXFLAGS[@]}" -fno-access-control -c "$HERE/ref_driver.cpp" -o "$STAGE/ref_driver.o" &
pids+=($!)
g++ -std=c++20 -O2 -ffp-contract=off -fPIC -c "$HERE/ref_table.cpp" -o "$STAGE/ref_table.o" &
pids+=($!)
gcc -O2 -ffp-contract=off -fPIC -I"$HERE/.." -c "$HERE/../rgb2spec.c" -o "$STAGE/rgb2spec.o" &
pids+=($!)
for p in "${pids[@]}"; do wait "$p"; done
g++ -shared -fopenmp -o "$OUT/libsrt_ref_${ORDER}${SUFFIX}.so" "${OBJS[@]}" "$STAGE/ref_driver.o" "$STAGE/ref_table.o" "$STAGE/rgb2spec.o" -lm
echo "build_ref.sh: built $OUT/libsrt_ref_${ORDER}${SUFFIX}.so"
