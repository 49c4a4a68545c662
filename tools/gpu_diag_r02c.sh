#!/bin/bash
# Diagnostic run: the layer-wise training gate twice (per-layer printout), the fixed / new tests of this session.
mkdir -p gpurun_out
for i in 1 2; do
  timeout 200 python -m pytest "tests/test_gpu_baseline_sizes.py::test_train_step_layerwise" -q -s -m gpu -p no:cacheprovider > gpurun_out/diag_layerwise_$i.log 2>&1
  echo "layerwise run $i rc=$?"
  grep -n "layer-wise on the engine\|^  \[layer" gpurun_out/diag_layerwise_$i.log | cut -c1-400
done
timeout 200 python -m pytest tests/test_gpu_kd.py tests/test_gpu_preprocess.py "tests/test_gpu_nms.py::test_tp_matching_matches_oracle" -q -s -m gpu -p no:cacheprovider > gpurun_out/diag_new.log 2>&1
echo "new tests rc=$?"; tail -n 5 gpurun_out/diag_new.log
