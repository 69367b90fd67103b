#!/bin/bash
# One GPU-box call: GPU tests, bench (ours + reference arm), ncu launch list, ncu captures of the three kernel groups.
# usage: tools/gpu_baseline.sh <tag> [skip_tests]
TAG=${1:-r2}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.txt 2>&1
if [ -z "$2" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_gputest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_gputest.log
  tail -3 $O/${TAG}_gputest.log
fi
timeout 900 python bench.py > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_reference_arm.json 2> $O/${TAG}_bench_ref.err; echo "ref rc=$?"
# launch list of a short bench run
CMD="python bench.py --steps 2 --warmup 3 --no-ref-cuda --no-cpu-baseline --no-extra"
timeout 300 $CMD > $O/${TAG}_plain_bench.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${TAG}_launches.csv $CMD > $O/${TAG}_ncu_launches.log 2>&1
echo "launch list rc=$?"
# top kernel: both launches (rounds) of one warm C2 render
CMD="python tools/perf_probe.py --variants wave"
timeout 300 $CMD > $O/${TAG}_plain_wave.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_wavefront -s 2 -c 2 -f -o $O/${TAG}_k_wavefront $CMD > $O/${TAG}_ncu_wave.log 2>&1
echo "wave ncu rc=$?"
cat $O/${TAG}_plain_wave.log
# LBVH build at 1M: the nine launches of one warm build
CMD="python tools/lbvh_only.py"
timeout 300 $CMD > $O/${TAG}_plain_lbvh.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -s 18 -c 9 -f -o $O/${TAG}_lbvh_1m $CMD > $O/${TAG}_ncu_lbvh.log 2>&1
echo "lbvh ncu rc=$?"
cat $O/${TAG}_plain_lbvh.log
# soup queries of the bench (primary + secondary rays at 1M and 10M): DRAM traffic per launch
CMD="python tools/lbvh_probe.py"
timeout 300 $CMD > $O/${TAG}_plain_trace.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none -k regex:'k_trace_rays|k_bounds|k_morton|k_onesweep|k_permute|k_build_tree|k_collapse4' -c 200 --csv --log-file $O/${TAG}_soup_traffic.csv $CMD > $O/${TAG}_ncu_trace.log 2>&1
echo "trace ncu rc=$?"
cat $O/${TAG}_plain_trace.log
ls -la $O | tail -30
