#!/bin/bash
# Run on the GPU box (gpurun) from the repo root. Writes into gpurun_out/.
export B2Y_NO_GRAPH=1      # eager launches so that every kernel is an individual ncu record
# (1) launch list: skip the start-up (input conversion, weight packing, 2 warm-up forwards), record ~2 forwards
ncu --metrics gpu__time_duration.sum --clock-control none -s 240 -c 260 --csv \
    --log-file gpurun_out/launches_r01.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench1.log 2>&1
# (2) full capture of the dominant kernel family: skip 3 forwards (75 conv_tc launches each: stem + 74 convs),
#     then record the first 16 conv_tc launches of a forward (stem, 320^2 / 160^2 / first 80^2 layers)
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 225 -c 16 \
    -o gpurun_out/prof_conv_early_r01 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench2.log 2>&1
# (3) the deep layers: 40x40 and 20x20 stages (launch index 27.. of a forward)
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 252 -c 12 \
    -o gpurun_out/prof_conv_deep_r01 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench3.log 2>&1
ls -la gpurun_out/*.ncu-rep
