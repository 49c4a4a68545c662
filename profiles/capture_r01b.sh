#!/bin/bash
# Round-1 (late) capture: pair-mode / weight-stationary conv kernel, 1x1 layers, head conv, decode.
export B2Y_NO_GRAPH=1
mkdir -p gpurun_out /tmp/ncu
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 222 -c 11 \
    -o /tmp/ncu/front -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 295 -c 1 \
    -o /tmp/ncu/head -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"yolo_decode|stem_fused" -s 8 -c 4 \
    -o /tmp/ncu/misc -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b3.log 2>&1
for n in front head misc; do
  ncu -i /tmp/ncu/$n.ncu-rep --page raw --csv > gpurun_out/ncu_${n}_raw_r01b.csv 2>/dev/null
  ncu -i /tmp/ncu/$n.ncu-rep --page details --csv > gpurun_out/ncu_${n}_details_r01b.csv 2>/dev/null
done
ncu -i /tmp/ncu/front.ncu-rep --page source --csv --kernel-id :::10 > gpurun_out/ncu_L013_source_r01b.csv 2>/dev/null
ncu -i /tmp/ncu/front.ncu-rep --page source --csv --kernel-id :::11 > gpurun_out/ncu_L014_source_r01b.csv 2>/dev/null
ncu -i /tmp/ncu/front.ncu-rep --page source --csv --kernel-id :::3 > gpurun_out/ncu_L003_source_r01b.csv 2>/dev/null
ncu -i /tmp/ncu/head.ncu-rep --page source --csv --kernel-id :::1 > gpurun_out/ncu_L105_source_r01b.csv 2>/dev/null
ls -la gpurun_out | tail -12
