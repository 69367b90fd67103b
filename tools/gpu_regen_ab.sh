#!/bin/bash
# One GPU-box call for the in-task sample regeneration (SRT_OPT_SCHED_FLAGS 256 / bits 12-25) and the banded whole-frame resolve:
# parity subset, A/B of the regeneration limits on C2 (rank 0 of 1 and of 8) and Prism, banded vs chunked resolve, a short bench line.
O=gpurun_out; mkdir -p $O
TAG=${1:-r2h}
timeout 600 python -m pytest tests -m gpu -x -q -k "rounds or pipelines or paths_in_flight or kernel_builds or c1_image or full_bench_size or whole_image or chunked or edge_cases or full_size" > $O/${TAG}_parity.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_parity.log
tail -3 $O/${TAG}_parity.log
FL="256,0,16793600,16842752,17039360,1081344,29392896,1310720,33816576"
timeout 300 python tools/ab_probe.py 1,8 0 $FL 2>&1 | tee $O/${TAG}_regen_ab.txt
timeout 300 python tools/ab_probe.py 1 1 256,0,17039360,1310720 2>&1 | tee -a $O/${TAG}_regen_ab.txt
timeout 300 python tools/ab_probe.py 1,8 0 256,0 2>&1 | tee -a $O/${TAG}_regen_ab.txt
timeout 120 python - <<'PY' 2>&1 | tee $O/${TAG}_resolve_check.txt
import sys, pathlib, numpy as np
sys.path.insert(0, "cuda-spectral-ray-tracer_b200")
import srt_b200 as S
for (w, h) in ((1920, 1080), (3840, 2160), (1000, 701)):
    a, xa, _ = S.render(scene_id=0, w=w, h=h, spp=2, bounce=10)                      # whole frame: banded resolve
    b, xb, _ = S.render(scene_id=0, w=w, h=h, spp=2, bounce=10, chunk=(256, 128))   # small chunks: one resolve per chunk
    print(w, h, "rgb equal:", bool(np.array_equal(a, b)), "xyz equal:", bool(np.array_equal(xa, xb)), "rgb sum", float(a.sum()))
PY
timeout 600 python bench.py --steps 5 --warmup 3 --no-ref-cuda --no-cpu-baseline --no-extra > $O/${TAG}_bench_short.json 2> $O/${TAG}_bench_short.err; echo "bench rc=$?"
python - <<PY
import json
d = json.load(open("$O/${TAG}_bench_short.json"))
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"], d["e2e_breakdown"], d["per_step"], "crc", d["film_crc"])
PY
