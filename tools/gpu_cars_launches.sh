mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_cars.csv python tools/bench_models.py --models cars --steps 1 --warmup 1 > gpurun_out/ncu_cars.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_duet.csv python tools/bench_models.py --models duet --steps 1 --warmup 1 > gpurun_out/ncu_duet.log 2>&1
tail -2 gpurun_out/ncu_cars.log | cut -c1-300
