# Late round-2 refresh after the training-forward / BPTT-grid / gemm_tc repack changes: training tests + timing, the captures whose
# source files changed (gemm_tc.cu users, training kernels), the training launch list.  ncu-rep files stay in /tmp.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_gpu.py -q -x 2>&1 | tail -3 > gpurun_out/pytest_train.log
timeout 300 python tools/train_timing.py 128 2>/dev/null | tail -1 > gpurun_out/train_timing.log
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc2_kernel|rnn_tc_kernel" -s 14 -c 10 -o /tmp/prof_r02f_cars -f python tools/bench_models.py --models cars --steps 1 --warmup 1 > gpurun_out/ncu_cars.log 2>&1
ncu -i /tmp/prof_r02f_cars.ncu-rep --page raw --csv > gpurun_out/r02_final_cars_gemm_rnn_ncu_raw.csv 2>/dev/null
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"lstm_tc_kernel|lstm_train_bwd_kernel|gemm_tn_kernel|mt_train_interact_bwd_kernel|embed_grad_kernel|mt_tc_interact_kernel" -s 30 -c 12 -o /tmp/prof_r02f_train -f python tools/train_timing.py 128 > gpurun_out/ncu_train.log 2>&1
ncu -i /tmp/prof_r02f_train.ncu-rep --page raw --csv > gpurun_out/r02_final_train_ncu_raw.csv 2>/dev/null
bash tools/gpu_train_launches.sh > gpurun_out/r02_final_train_launches.txt 2>&1
cat gpurun_out/pytest_train.log gpurun_out/train_timing.log; head -16 gpurun_out/r02_final_train_launches.txt
