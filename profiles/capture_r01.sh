#!/bin/bash
# Run on the GPU box (gpurun) from the repo root. Writes small CSV summaries into gpurun_out/ (the raw .ncu-rep files
# exceed the 64 MiB return limit, so they are exported to CSV on the box and removed).
export B2Y_NO_GRAPH=1      # eager launches so that every kernel is an individual ncu record
mkdir -p gpurun_out /tmp/ncu
# (1) launch list: skip the start-up (input conversion, weight packing, 2 warm-up forwards), record ~2 forwards
ncu --metrics gpu__time_duration.sum --clock-control none -s 240 -c 260 --csv \
    --log-file gpurun_out/launches_r01.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench1.log 2>&1
# (2) full capture of the dominant kernel family (conv_tc_kernel): skip 3 forwards (75 launches each), then
#     launches 0..13 of a forward = stem, the 320^2 / 160^2 stages and the first 80^2 layers
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 225 -c 14 \
    -o /tmp/ncu/early -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench2.log 2>&1
# (3) deep layers: 40x40 (launch 27..) and 20x20 (launch 44..)
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 252 -c 3 \
    -o /tmp/ncu/deep40 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 269 -c 3 \
    -o /tmp/ncu/deep20 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench4.log 2>&1
for n in early deep40 deep20; do
  ncu -i /tmp/ncu/$n.ncu-rep --page raw --csv > gpurun_out/ncu_${n}_raw_r01.csv 2>/dev/null
  ncu -i /tmp/ncu/$n.ncu-rep --page details --csv > gpurun_out/ncu_${n}_details_r01.csv 2>/dev/null
done
# per-instruction stall sampling of one representative kernel (3x3 128->256 @80x80 + residual = launch 11)
ncu -i /tmp/ncu/early.ncu-rep --page source --csv --kernel-id :::12 > gpurun_out/ncu_L014_source_r01.csv 2>/dev/null
ls -la /tmp/ncu gpurun_out | tail -20
cp /tmp/ncu/deep40.ncu-rep gpurun_out/prof_conv_deep40_r01.ncu-rep 2>/dev/null
du -sh gpurun_out
