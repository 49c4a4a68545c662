#!/bin/bash
# Launch list (durations only) of two eager YOLOv4 training steps, 8 images, 640x640.
export B2Y_NO_GRAPH=1
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/launches_train_r01.csv \
    python tools/bench_train.py --steps 1 --warmup 1 > gpurun_out/ncu_t1.log 2>&1
tail -2 gpurun_out/ncu_t1.log; wc -l gpurun_out/launches_train_r01.csv
