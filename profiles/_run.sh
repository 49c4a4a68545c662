cd /root/repo
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | grep -v "Warning\|^$" | tail -6
timeout 600 python tools/bench_train.py --steps 10 2>/dev/null | tail -1
timeout 600 python tools/bench_train.py --model yolov3 --steps 10 2>/dev/null | tail -1
