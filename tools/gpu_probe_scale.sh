#!/bin/bash
# single-GPU proxies of the N-rank split; $1 = tag, $2 = flags list, $3 = worlds
O=gpurun_out; mkdir -p $O; T=${1:-x}
python tools/sched_probe.py ${3:-1,2,4,8} ${2:-0} 0 > $O/p_sched_$T.log 2>&1
python tools/scale_probe_c5.py > $O/p_scale_c5_$T.log 2>&1
cat $O/p_sched_$T.log; head -2 $O/p_scale_c5_$T.log
