cd /root/repo
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x 2>&1 | tail -12
timeout 300 python -m pytest tests/test_gpu_model.py -m gpu -q -x 2>&1 | tail -5
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --profile-layers gpurun_out/layers_pair.json 2>gpurun_out/bench12.err | head -c 250; echo; tail -3 gpurun_out/bench12.err
B2Y_PAIR=0 B2Y_CLUSTER=1 timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --profile-layers gpurun_out/layers_nopair.json 2>/dev/null | head -c 250; echo
