cd /root/repo
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py tests/test_gpu_train_kernels.py -m gpu -q -x 2>&1 | tail -3
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --profile-layers gpurun_out/layers_t0.json 2>/dev/null | head -c 200; echo
B2Y_EPI_TMA=0 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --profile-layers gpurun_out/layers_t1.json 2>/dev/null | head -c 200; echo
for d in 1 2 3; do B2Y_EPI_DEBUG=$d timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-layers gpurun_out/layers_d$d.json 2>/dev/null | head -c 200; echo; done
