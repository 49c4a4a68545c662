cd /root/repo
timeout 900 python -m pytest tests/test_gpu_train_model.py -m gpu -q -x -s 2>&1 | grep -v "Warning\|^$" | tail -24
