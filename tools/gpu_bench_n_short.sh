#!/bin/bash
# tools/gpu_bench_n_short.sh <N> <tag>: N-rank bench line without the CPU / reference-CUDA legs (gpurun --gpus N), JSON line only
N=$1; T=${2:-r2}; O=gpurun_out; mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-ref-cuda --no-cpu-baseline > $O/${T}_bench_n$N.json 2> $O/${T}_bench_n$N.err
echo "rc=$?"
python - <<PY
import json
txt = open("$O/${T}_bench_n$N.json").read()
d = json.loads([l for l in txt.splitlines() if l.startswith("{")][-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"], d["e2e_breakdown"], d["per_step"], "crc", d["film_crc"])
for k in ("c3", "c5", "soup"): print(k, d[k] and d[k]["value"])
PY
