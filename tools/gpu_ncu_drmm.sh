mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:drmm_tc_kernel -s 2 -c 1 -o gpurun_out/prof_drmm_tc -f python tools/bench_models.py --models drmm --steps 1 --warmup 2 > gpurun_out/ncu_drmm.log 2>&1
tail -2 gpurun_out/ncu_drmm.log
ncu -i gpurun_out/prof_drmm_tc.ncu-rep --page raw --csv > gpurun_out/prof_drmm_tc_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_drmm_tc.ncu-rep --page source --csv > gpurun_out/prof_drmm_tc_source.csv 2>/dev/null
python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/prof_drmm_tc_raw.csv')))
hdr=rows[0]
want=['Grid Size','gpu__time_duration.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','dram__bytes_read.sum','dram__bytes_write.sum','lts__throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','sm__throughput.avg.pct_of_peak_sustained_elapsed','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','launch__occupancy_limit_shared_mem','sm__cycles_elapsed.max','launch__waves_per_multiprocessor','sm__inst_executed.sum','smsp__inst_executed.avg.per_cycle_active','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum']
for r in rows[2:]:
    print({w: r[hdr.index(w)] for w in want if w in hdr})
rows=list(csv.reader(open('gpurun_out/prof_drmm_tc_source.csv')))
hdr=rows[1]; ci=hdr.index('Source'); si=hdr.index('# Samples')
stalls=[i for i,h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
data=[]
for r in rows[2:]:
    try: data.append((int(r[si]), r))
    except: pass
tot=sum(d[0] for d in data); print('total samples',tot)
data.sort(key=lambda d:-d[0])
for n,r in data[:22]:
    top=sorted(((int(r[i] or 0),hdr[i]) for i in stalls), reverse=True)[:2]
    print('%6d %5.1f%%  %-70s %s' % (n, 100*n/tot, r[ci][:70], top))
PY
