cd /root/repo
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --profile-layers gpurun_out/layers_h1.json > gpurun_out/bench17.json 2>gpurun_out/bench17.err; tail -2 gpurun_out/bench17.err; head -c 1500 gpurun_out/bench17.json
