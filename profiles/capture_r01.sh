#!/bin/bash
# Run on the GPU box (gpurun) from the repo root. Writes into gpurun_out/.
set -x
export B2Y_NO_GRAPH=1      # eager launches so that every kernel is an individual ncu record
# (1) launch list of one forward: skip weight packing (77) + 3 warm-up forwards (83 each)
ncu --metrics gpu__time_duration.sum --clock-control none -s 326 -c 83 --csv \
    --log-file gpurun_out/launches_r01.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench1.log 2>&1
# (2) full capture of the dominant kernel: the 80x80 stage of Darknet-53 (3x3 128->256 + residual, 1x1 256->128)
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 236 -c 4 \
    -o gpurun_out/prof_conv_r01 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench2.log 2>&1
# (3) one deep 3x3 (256->512 @40x40 + residual)
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 254 -c 2 \
    -o gpurun_out/prof_conv40_r01 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench3.log 2>&1
ls -la gpurun_out/*.ncu-rep
