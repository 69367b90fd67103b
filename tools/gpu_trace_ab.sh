#!/bin/bash
# soup queries (primary / secondary rays) with several builds of the library: tools/gpu_trace_ab.sh <triangle counts> <lib>...
N=$1; shift
for L in "$@"; do echo "== $L"; SRT_LIB=$PWD/cuda-spectral-ray-tracer_b200/$L timeout 600 python tools/lbvh_probe.py $N 2>&1 | grep -E "primary|secondary|render"; done
