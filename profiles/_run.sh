cd /root/repo
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "stem" 2>&1 | tail -4
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --profile-layers gpurun_out/layers_r01j.json > gpurun_out/bench10.json 2> gpurun_out/bench10.err; tail -3 gpurun_out/bench10.err; head -c 250 gpurun_out/bench10.json; echo
B2Y_STEM_NG=2 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --profile-layers gpurun_out/layers_r01k.json > gpurun_out/bench10_ng2.json 2>/dev/null; head -c 250 gpurun_out/bench10_ng2.json; echo
