#!/bin/bash
# A/B of library builds / scheduler flags on one box: tools/gpu_ab.sh "<worlds>" "<flags>" libA.so libB.so ...   (C2 and Prism frames, rank 0 of world N)
O=gpurun_out; mkdir -p $O
W=$1; F=$2; shift; shift
for rep in 1 2; do
for L in "$@"; do
  SRT_LIB=$PWD/cuda-spectral-ray-tracer_b200/$L timeout 300 python tools/ab_probe.py $W 0 $F 2>&1 | tee -a $O/p_ab.log
done
done
for L in "$@"; do
  SRT_LIB=$PWD/cuda-spectral-ray-tracer_b200/$L timeout 300 python tools/ab_probe.py 1 1 $F 2>&1 | tee -a $O/p_ab.log
done
