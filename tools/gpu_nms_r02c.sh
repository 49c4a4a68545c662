#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_nms.py -q -s -m gpu -p no:cacheprovider > gpurun_out/nms_opt_tests.log 2>&1
echo "nms tests rc=$?"; tail -n 3 gpurun_out/nms_opt_tests.log
timeout 100 python tools/nms_profile.py > gpurun_out/nms_opt_timing.log 2>&1; echo "timing rc=$?"; tail -n 1 gpurun_out/nms_opt_timing.log | cut -c1-300
timeout 120 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/launches_nms_opt.csv python tools/nms_profile.py > /dev/null 2>&1; echo "ncu rc=$?"
grep -c "nms_" gpurun_out/launches_nms_opt.csv
