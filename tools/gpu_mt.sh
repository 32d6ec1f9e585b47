# Match-Tensor-focused pass: MT parity tests, short bench, interact role timing. logs under gpurun_out/
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "match_tensor or selftest" 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench.log 2>&1
tail -1 gpurun_out/bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['stages_ms'])"
timeout 300 python tools/mt_timing.py > gpurun_out/mt_timing.log 2>&1
tail -12 gpurun_out/mt_timing.log
