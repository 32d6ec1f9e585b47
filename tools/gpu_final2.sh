# Round-2 final evidence: GPU tests, smoke, bench (both arms), launch list, ncu --set full of every kernel family, timing tools.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 900 python bench.py > gpurun_out/bench.log 2>&1
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1
# launch list of two headline steps (cold-cache, serialised: compare SHARES with the live stage split)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_final_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-other-configs > gpurun_out/ncu_launches.log 2>&1
# every kernel of the headline step
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"lstm_tc_kernel|mt_tc_interact_kernel|mt_tc_proj_image_kernel|mt_tc_build_t_kernel|gemm_f32_kernel" -s 12 -c 6 -o /tmp/prof_r02f_cfg2 -f python tools/one_batch.py > gpurun_out/ncu_cfg2.log 2>&1
ncu -i /tmp/prof_r02f_cfg2.ncu-rep --page raw --csv > gpurun_out/r02_final_cfg2_step_ncu_raw.csv 2>/dev/null
# CARS: persistent GEMM (pre-gates), cluster recurrence
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc2_kernel|rnn_tc_kernel" -s 14 -c 10 -o /tmp/prof_r02f_cars -f python tools/bench_models.py --models cars --steps 1 --warmup 1 > gpurun_out/ncu_cars.log 2>&1
ncu -i /tmp/prof_r02f_cars.ncu-rep --page raw --csv > gpurun_out/r02_final_cars_gemm_rnn_ncu_raw.csv 2>/dev/null
# DRMM tcgen05 kernel
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"drmm_tc_kernel" -s 1 -c 1 -o /tmp/prof_r02f_drmm -f python tools/bench_models.py --models drmm --steps 1 --warmup 1 > gpurun_out/ncu_drmm.log 2>&1
ncu -i /tmp/prof_r02f_drmm.ncu-rep --page raw --csv > gpurun_out/r02_final_drmm_tc_ncu_raw.csv 2>/dev/null
# training step kernels
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"mt_interact_kernel|lstm_train_bwd_kernel|lstm_train_fwd_kernel|gemm_tn_kernel|mt_train_interact_bwd_kernel|embed_grad_kernel" -s 40 -c 12 -o /tmp/prof_r02f_train -f python tools/train_timing.py 128 > gpurun_out/ncu_train.log 2>&1
ncu -i /tmp/prof_r02f_train.ncu-rep --page raw --csv > gpurun_out/r02_final_train_ncu_raw.csv 2>/dev/null
timeout 300 python tools/train_timing.py 128 > gpurun_out/train_timing.log 2>&1
timeout 300 python tools/lstm_timing.py > gpurun_out/lstm_timing.log 2>&1
timeout 300 python tools/drmm_timing.py > gpurun_out/drmm_timing.log 2>&1
timeout 300 python tools/gemm2_experiments.py > gpurun_out/gemm2_experiments.log 2>&1
timeout 300 python tools/umma_bench.py > gpurun_out/umma_bench.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; tail -1 gpurun_out/bench.log | cut -c1-600; tail -1 gpurun_out/bench_ref.log | cut -c1-400
python - <<'PY'
import csv
for f in ('r02_final_cfg2_step_ncu_raw.csv','r02_final_cars_gemm_rnn_ncu_raw.csv','r02_final_drmm_tc_ncu_raw.csv','r02_final_train_ncu_raw.csv'):
    try:
        rows=list(csv.reader(open('gpurun_out/'+f)))
    except Exception as e:
        print(f, e); continue
    hdr=rows[0]
    want=['Kernel Name','launch__grid_size','gpu__time_duration.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed','dram__bytes_read.sum','dram__bytes_write.sum','launch__registers_per_thread']
    print(f)
    for r in rows[2:]:
        print('  ', [r[hdr.index(w)][:44] for w in want if w in hdr])
PY
