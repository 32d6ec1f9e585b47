mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "cars or metrics or duet or arc or dssm" 2>&1 | tail -25 > gpurun_out/pytest_cars.log
tail -12 gpurun_out/pytest_cars.log
timeout 600 python tools/bench_models.py --models cars,duet --steps 5 > gpurun_out/bench_models.log 2>&1
cut -c1-600 gpurun_out/bench_models.log
