cd /root/repo
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -m gpu -q -x 2>&1 | tail -4
timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>gpurun_out/bench13.err | head -c 250; echo; tail -3 gpurun_out/bench13.err
