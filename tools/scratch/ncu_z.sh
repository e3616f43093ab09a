#!/bin/bash
ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k 'regex:pose_fused_kernel<\(int\)0, \(int\)[01],' -c 6 -o gpurun_out/prof_xyz -f python tools/bench_ops.py --fast-only --reps 1 > gpurun_out/ncu_xyz.log 2>&1
tail -3 gpurun_out/ncu_xyz.log
