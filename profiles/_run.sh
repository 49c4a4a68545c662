cd /root/repo
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py tests/test_gpu_train_kernels.py -m gpu -q -x 2>&1 | tail -4
timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>gpurun_out/bench14.err | head -c 250; echo; tail -3 gpurun_out/bench14.err
B2Y_PDL=0 timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | head -c 250; echo
timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | head -c 250; echo
