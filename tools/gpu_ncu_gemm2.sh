mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc2_kernel -s 12 -c 8 -o gpurun_out/prof_gemm2_cars -f python tools/bench_models.py --models cars --steps 2 > gpurun_out/ncu_gemm2.log 2>&1
ncu -i gpurun_out/prof_gemm2_cars.ncu-rep --page raw --csv > gpurun_out/prof_gemm2_cars_raw.csv 2>/dev/null
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/prof_gemm2_cars_raw.csv')))
hdr=rows[0]
want=['launch__grid_size','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed','smsp__issue_active.avg.pct','lts__t_sector_hit_rate.pct','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','launch__registers_per_thread']
for r in rows[2:]:
    d=dict(zip(hdr,r))
    print({w.split('.')[0][-28:]: d.get(w) for w in want})
    st={k.replace('smsp__pcsamp_warps_issue_stalled_',''):int(d[k]) for k in d if k.startswith('smsp__pcsamp_warps_issue_stalled_') and not k.endswith('not_issued') and d[k] not in ('','0')}
    print('   stalls', sorted(st.items(), key=lambda x:-x[1])[:8])
PY
