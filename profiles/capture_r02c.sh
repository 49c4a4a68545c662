#!/bin/bash
# End-of-round-2 evidence (third session): full GPU suite with its printed parity values, the bench line, an ncu launch list
# of the NMS path.  Run on the GPU box:  bash profiles/capture_r02c.sh   (results land in gpurun_out/, copies in profiles/r02c/)
mkdir -p gpurun_out
S=gpurun_out/summary_final_r02c.txt
: > $S
timeout 420 python -m pytest tests -q -m gpu -s -p no:cacheprovider > gpurun_out/pytest_gpu_final_r02c.log 2>&1
echo "pytest rc=$?" >> $S
tail -n 4 gpurun_out/pytest_gpu_final_r02c.log >> $S
timeout 300 python bench.py > gpurun_out/bench_n1_final_r02c.json 2> gpurun_out/bench_n1_final_r02c.err
echo "bench rc=$?" >> $S
timeout 150 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/launches_nms_r02c.csv python tools/nms_profile.py > gpurun_out/nms_profile.log 2>&1
echo "ncu rc=$?" >> $S
cat $S
