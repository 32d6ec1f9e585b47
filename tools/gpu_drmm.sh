mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "drmm" 2>&1 | tail -25 > gpurun_out/pytest_drmm.log
tail -6 gpurun_out/pytest_drmm.log
timeout 300 python tools/drmm_timing.py 2>&1 | tail -28
