mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 3 -c 3 -o gpurun_out/prof_gemm_tc -f python tools/bench_models.py --models duet --steps 1 --warmup 1 > gpurun_out/ncu_gemm.log 2>&1
tail -3 gpurun_out/ncu_gemm.log
ncu -i gpurun_out/prof_gemm_tc.ncu-rep --page raw --csv > gpurun_out/prof_gemm_tc_raw.csv 2>/dev/null
python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/prof_gemm_tc_raw.csv')))
hdr=rows[0]
want=['Kernel Name','Grid Size','gpu__time_duration.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_tensor.sum','dram__bytes_read.sum','dram__bytes_write.sum','lts__t_bytes.sum','lts__throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','l1tex__t_bytes.sum','sm__throughput.avg.pct_of_peak_sustained_elapsed','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','launch__occupancy_limit_shared_mem','sm__cycles_elapsed.max']
idx=[hdr.index(w) for w in want if w in hdr]
for r in rows[2:]:
    print({hdr[i]: r[i] for i in idx})
PY
