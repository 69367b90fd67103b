#!/bin/bash
O=gpurun_out; mkdir -p $O
export SRT_LIB=$PWD/cuda-spectral-ray-tracer_b200/${1:-libsrt_prof.so}
python tools/latency_probe.py > $O/p_latency_${2:-a}.log 2>&1
cat $O/p_latency_${2:-a}.log
