#!/bin/bash
# Per-instruction stall sampling of the two small-K layers that dominate the front of YOLOv3 (stem GEMM, L001 3x3 s2 32->64).
export B2Y_NO_GRAPH=1
mkdir -p gpurun_out /tmp/ncu
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 225 -c 3 \
    -o /tmp/ncu/smallk -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench5.log 2>&1
ncu -i /tmp/ncu/smallk.ncu-rep --page source --csv --kernel-id :::1 > gpurun_out/ncu_stem_source_r01.csv 2>/dev/null
ncu -i /tmp/ncu/smallk.ncu-rep --page source --csv --kernel-id :::2 > gpurun_out/ncu_L001_source_r01.csv 2>/dev/null
ncu -i /tmp/ncu/smallk.ncu-rep --page source --csv --kernel-id :::3 > gpurun_out/ncu_L002_source_r01.csv 2>/dev/null
ncu -i /tmp/ncu/smallk.ncu-rep --page raw --csv > gpurun_out/ncu_smallk_raw_r01.csv 2>/dev/null
ls -la gpurun_out | tail -8
