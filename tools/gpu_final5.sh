# Last refresh of round 2: full GPU suite, smoke, bench (both arms), launch list, the captures whose sources changed (gemm_tc.cu users).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 900 python bench.py > gpurun_out/bench.log 2>&1
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_final_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-other-configs > gpurun_out/ncu_launches.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc2_kernel|gemm_tc_aimg_kernel|rnn_tc_kernel" -s 14 -c 12 -o /tmp/prof_r02f_cars -f python tools/bench_models.py --models cars --steps 1 --warmup 1 > gpurun_out/ncu_cars.log 2>&1
ncu -i /tmp/prof_r02f_cars.ncu-rep --page raw --csv > gpurun_out/r02_final_cars_gemm_rnn_ncu_raw.csv 2>/dev/null
timeout 900 compute-sanitizer --tool memcheck --report-api-errors no --error-exitcode 3 --print-limit 20 python -m pytest tests/test_parity_gpu.py tests/test_session_rankers.py -q -x -m gpu -k "cars_golden or duet_golden or mnsrf or match_tensor_golden" > gpurun_out/sanitize_aimg.log 2>&1
echo "rc=$?" >> gpurun_out/sanitize_aimg.log
cat gpurun_out/pytest_gpu.log; tail -1 gpurun_out/smoke.log; tail -3 gpurun_out/sanitize_aimg.log
