# Final refresh of the headline evidence with the last code: tests, smoke, both bench arms, launch list, ncu of the cfg2 step.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 900 python bench.py > gpurun_out/bench.log 2>&1
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_final_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-other-configs > gpurun_out/ncu_launches.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"lstm_tc_kernel|mt_tc_interact_kernel|mt_tc_proj_image_kernel|mt_tc_build_t_kernel|gemm_f32_kernel" -s 12 -c 6 -o /tmp/prof_r02f_cfg2 -f python tools/one_batch.py > gpurun_out/ncu_cfg2.log 2>&1
ncu -i /tmp/prof_r02f_cfg2.ncu-rep --page raw --csv > gpurun_out/r02_final_cfg2_step_ncu_raw.csv 2>/dev/null
timeout 300 python tools/mt_realistic_probe.py > gpurun_out/mt_realistic_probe.log 2>&1
timeout 300 python tools/mt_timing.py > gpurun_out/mt_timing.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; tail -1 gpurun_out/smoke.log
