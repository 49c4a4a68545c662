#!/bin/bash
# Launch list with DRAM bytes for every kernel of the forward (eager launches), final round-1 kernels.
export B2Y_NO_GRAPH=1
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -k regex:"conv_tc_kernel|stem_fused|yolo_decode|upsample" -c 330 --csv --log-file gpurun_out/launches_r01c.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c1.log 2>&1
tail -2 gpurun_out/ncu_c1.log; wc -l gpurun_out/launches_r01c.csv
