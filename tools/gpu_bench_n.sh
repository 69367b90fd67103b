#!/bin/bash
# tools/gpu_bench_n.sh <N> <tag>: the driver's launch line for N ranks on one box (gpurun --gpus N)
N=$1; T=${2:-r2}; O=gpurun_out; mkdir -p $O
if [ "$N" = "1" ]; then
  timeout 900 python bench.py --gpus 1 > $O/${T}_bench_n1.json 2> $O/${T}_bench_n1.err
else
  NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > $O/${T}_bench_n$N.json 2> $O/${T}_bench_n$N.err
fi
echo "rc=$?"; tail -c 3000 $O/${T}_bench_n$N.json; grep -E "NCCL INFO (comm|Connected|NVLS|ncclCommInitRank)" $O/${T}_bench_n$N.err | head -12; tail -3 $O/${T}_bench_n$N.err
