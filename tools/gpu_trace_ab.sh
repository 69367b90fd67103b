#!/bin/bash
# soup queries (primary / secondary rays at 1M and 10M triangles) with several builds of the library
for L in "$@"; do echo "== $L"; SRT_LIB=$PWD/cuda-spectral-ray-tracer_b200/$L timeout 600 python tools/lbvh_probe.py 2>&1 | grep -E "primary|secondary|render"; done
