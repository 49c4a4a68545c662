#!/bin/bash
# Launch list (durations) of the INT8 PTQ forward, batch 32, 640x640.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_ptq_r01.csv \
    python tests/bench_ptq.py --steps 1 --warmup 1 > gpurun_out/ncu_p1.log 2>&1
tail -1 gpurun_out/ncu_p1.log | cut -c1-200; wc -l gpurun_out/launches_ptq_r01.csv
