# pipeline-focused pass: full-shape property test (incl. submit/wait pipeline), bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "full_cfg2 or golden" 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench.log 2>&1
tail -1 gpurun_out/bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e'], d['stages_ms'])"
