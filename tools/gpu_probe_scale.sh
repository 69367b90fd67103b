#!/bin/bash
O=gpurun_out; mkdir -p $O
python tools/scale_probe.py > $O/p_scale_c2.log 2>&1
python tools/pass_log_probe.py 8 1 > $O/p_passlog_w8.log 2>&1
python tools/pass_log_probe.py 1 2 > $O/p_passlog_w1.log 2>&1
python tools/sweep_probe.py 8 0,128,256,512,1024 0 1,2 0 > $O/p_sweep_w8.log 2>&1
python tools/sweep_probe.py 4 0,256,512,1024 0 1,2 0 > $O/p_sweep_w4.log 2>&1
python tools/scale_probe_c5.py > $O/p_scale_c5.log 2>&1
tail -n 40 $O/p_scale_c2.log $O/p_sweep_w8.log $O/p_sweep_w4.log $O/p_scale_c5.log
