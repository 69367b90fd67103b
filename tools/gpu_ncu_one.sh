#!/bin/bash
# tools/gpu_ncu_one.sh <tag> <kernel regex> <skip> <count> <cmd...>: plain run, then one ncu --set full capture
O=gpurun_out; mkdir -p $O
T=$1; K=$2; SKIP=$3; CNT=$4; shift; shift; shift; shift
timeout 300 "$@" > $O/${T}_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c $CNT -f -o $O/$T "$@" > $O/${T}_ncu.log 2>&1
echo "rc=$?"; cat $O/${T}_plain.log; tail -3 $O/${T}_ncu.log
