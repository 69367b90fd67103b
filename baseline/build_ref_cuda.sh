#!/usr/bin/env bash
# BASELINE INFRASTRUCTURE: builds the reference's own CUDA renderer for sm_100a into
# baseline/_ref/ref_cuda_render (git-ignored, travels with gpurun).  Sources are compiled from a
# scratch copy of /root/reference staged outside the repository; the reference's CMake is not used
# (its arch list is 50;75;80, CMakeLists.txt:56).  Edits to the scratch copy:
#   * utils/device_init.cuh:34 removed: cudaMemcpyToSymbol(&symbol, ...) passes the ADDRESS of the
#     symbol -> cudaErrorInvalidSymbol -> exit(99); the symbol is never read.
# Flags: -std=c++20 -rdc=true (the `__constant__ inline` symbols need it) -include cfloat
# (math/interval.cuh:14 uses FLT_MAX without the header).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${SRT_REFERENCE_DIR:-/root/reference}"
OUT="$HERE/_ref"
mkdir -p "$OUT"
# library sort next to our onesweep (timing context only); needs no reference sources
nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -w "$HERE/cub_sort_main.cu" -o "$OUT/cub_sort_bench"
if [[ ! -d "$REF" ]]; then echo "build_ref_cuda.sh: $REF not present -- keeping prebuilt baseline/_ref" >&2; exit 0; fi
STAGE="$(mktemp -d /tmp/srt_refcuda_stage.XXXXXX)"
trap 'rm -rf "$STAGE"' EXIT
for d in materials primitives bvh utils rendering refraction color spectrum ray math io scene _log_; do cp -r "$REF/$d" "$STAGE/$d"; done
sed -i '/cudaMemcpyToSymbol(&dev_sRGBToSpectrumTable_Res/d' "$STAGE/utils/device_init.cuh"
INC=(-I"$STAGE")
for d in materials primitives bvh utils rendering refraction color spectrum ray math io scene _log_; do INC+=(-I"$STAGE/$d"); done
NV=(nvcc -std=c++20 -O3 -rdc=true -gencode arch=compute_100a,code=sm_100a -include cfloat -w "${INC[@]}")
SRCS=(utils/cie_const.cu spectrum/spectrum.cu utils/color_const.cu utils/cuda_utility.cu color/color.cu materials/material.cu bvh/aabb.cu
      bvh/bvh.cu rendering/rendering.cu rendering/camera.cu scene/scene.cu primitives/transform.cu refraction/sellmeier.cu primitives/tri.cu
      primitives/tri_quad.cu primitives/prism.cu primitives/tri_box.cu primitives/pyramid.cu rendering/render_manager.cu)
OBJS=(); pids=()
for s in "${SRCS[@]}"; do
  o="$STAGE/$(echo "$s" | tr '/.' '__').o"; OBJS+=("$o")
  "${NV[@]}" -c "$STAGE/$s" -o "$o" & pids+=($!)
done
"${NV[@]}" -c "$HERE/ref_cuda_main.cu" -o "$STAGE/main.o" & pids+=($!)
g++ -std=c++20 -O2 -c "$STAGE/_log_/log_context.cpp" -I"$STAGE/utils" -I"$STAGE/_log_" -o "$STAGE/log.o" & pids+=($!)
g++ -std=c++20 -O2 -c "$STAGE/io/params.cpp" -I"$STAGE/io" -o "$STAGE/params.o" & pids+=($!)
g++ -std=c++20 -O2 -ffp-contract=off -I"$HERE/../oracle" -c "$HERE/ref_table.cpp" -o "$STAGE/table.o" & pids+=($!)
gcc -O2 -ffp-contract=off -I"$HERE/../oracle" -c "$HERE/../oracle/rgb2spec.c" -o "$STAGE/rgb2spec.o" & pids+=($!)
for p in "${pids[@]}"; do wait "$p"; done
nvcc -rdc=true -gencode arch=compute_100a,code=sm_100a "${OBJS[@]}" "$STAGE/main.o" "$STAGE/log.o" "$STAGE/params.o" "$STAGE/table.o" "$STAGE/rgb2spec.o" -o "$OUT/ref_cuda_render"
echo "build_ref_cuda.sh: built $OUT/ref_cuda_render"
