#!/bin/bash
# One-box validation run of round 2 (third session): full GPU suite (prints kept), bench line, memcheck of the NMS path.
#   gpurun --timeout 900 -- 'bash tools/gpu_check_r02c.sh'
mkdir -p gpurun_out
S=gpurun_out/summary_r02c.txt
: > $S
date +%s >> $S
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv >> $S 2>&1
timeout 420 python -m pytest tests -q -m gpu -s -p no:cacheprovider > gpurun_out/pytest_gpu_r02c.log 2>&1
echo "pytest rc=$?" >> $S
tail -n 3 gpurun_out/pytest_gpu_r02c.log >> $S
date +%s >> $S
timeout 420 python bench.py > gpurun_out/bench_n1_r02c.json 2> gpurun_out/bench_n1_r02c.err
echo "bench rc=$?" >> $S
date +%s >> $S
timeout 200 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
    python -m pytest tests/test_gpu_nms.py -q -m gpu -p no:cacheprovider -k "fixture or chunk_edges or many_kept or tp_matching" \
    > gpurun_out/sanitizer_nms_r02c.log 2>&1
echo "sanitizer rc=$?" >> $S
tail -n 5 gpurun_out/sanitizer_nms_r02c.log >> $S
date +%s >> $S
cat $S
