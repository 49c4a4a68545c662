cd /root/repo
timeout 600 python -m pytest tests/test_gpu_ptq.py tests/test_gpu_train_kernels.py -m gpu -q -x -s 2>&1 | grep -v "Warning\|^$" | tail -6
timeout 600 python tools/bench_ptq.py --steps 10 2>/dev/null | tail -1
