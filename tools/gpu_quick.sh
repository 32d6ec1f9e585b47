# usage: bash tools/gpu_quick.sh [pytest -k expr]   -> GPU tests + short bench, logs under gpurun_out/
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x ${1:+-k "$1"} 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1
tail -3 gpurun_out/bench.log | cut -c1-3000
