# Round-end evidence: GPU tests, smoke, bench (both arms), ncu launch list, ncu --set full of the two top kernels.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
lscpu | head -20 > gpurun_out/lscpu.txt
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 600 python bench.py > gpurun_out/bench.log 2>&1
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mt_tc_interact -s 3 -c 1 -o gpurun_out/prof_interact -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_interact.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lstm_tc_kernel -s 6 -c 2 -o gpurun_out/prof_lstm -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_lstm.log 2>&1
timeout 300 python tools/lstm_timing.py > gpurun_out/lstm_timing.log 2>&1
timeout 300 python tools/mt_timing.py > gpurun_out/mt_timing.log 2>&1
timeout 600 python tools/bench_models.py > gpurun_out/bench_models.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; tail -1 gpurun_out/bench.log | cut -c1-1500; tail -1 gpurun_out/bench_ref.log | cut -c1-600; tail -12 gpurun_out/bench_models.log
