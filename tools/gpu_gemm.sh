mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "duet or cars or dssm or arc" 2>&1 | tail -5
timeout 600 python tools/bench_models.py --models duet,cars 2>&1 | tail -4 | cut -c1-300
