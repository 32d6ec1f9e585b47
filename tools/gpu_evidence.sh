mkdir -p gpurun_out
# 1. launch list of one bench step (cold-cache, serialised: compare shares)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-other-configs > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log | cut -c1-200
# 2. every kernel of the headline step, third forward
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"lstm_tc_kernel|rnn_tc_kernel|mt_tc_interact_kernel|mt_tc_proj_image_kernel|mt_tc_build_t_kernel|gemm_f32_kernel" -s 12 -c 6 -o gpurun_out/prof_r02_cfg2 -f python tools/one_batch.py > gpurun_out/ncu_cfg2.log 2>&1
tail -2 gpurun_out/ncu_cfg2.log
ncu -i gpurun_out/prof_r02_cfg2.ncu-rep --page raw --csv > gpurun_out/r02_cfg2_step_ncu_raw.csv 2>/dev/null
# 3. the cluster-split recurrence on the CARS document encoder (h = 128 / direction, 4-CTA clusters)
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"rnn_tc_kernel" -s 2 -c 2 -o gpurun_out/prof_r02_cars_rnn -f python tools/bench_models.py --models cars --steps 1 --warmup 1 > gpurun_out/ncu_cars.log 2>&1
tail -2 gpurun_out/ncu_cars.log | cut -c1-200
ncu -i gpurun_out/prof_r02_cars_rnn.ncu-rep --page raw --csv > gpurun_out/r02_cars_rnn_tc_ncu_raw.csv 2>/dev/null
python - <<PY
import csv
for f in ('gpurun_out/r02_cfg2_step_ncu_raw.csv','gpurun_out/r02_cars_rnn_tc_ncu_raw.csv'):
    rows=list(csv.reader(open(f)))
    hdr=rows[0]
    want=['Kernel Name','Grid Size','gpu__time_duration.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','dram__bytes_read.sum','dram__bytes_write.sum','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']
    for r in rows[2:]:
        print({w[:40]: r[hdr.index(w)][:60] for w in want if w in hdr})
PY
