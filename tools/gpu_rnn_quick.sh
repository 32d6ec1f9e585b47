mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x -k "lstm_entry or rnn_uni or match_tensor_golden or cars_golden" 2>&1 | tail -15 > gpurun_out/pytest_rnn.log
tail -5 gpurun_out/pytest_rnn.log
timeout 300 python tools/rnn_timing.py > gpurun_out/rnn_timing.log 2>&1
grep -v "epi warp  *[2-9]:\|epi warp 1[0-8]" gpurun_out/rnn_timing.log | tail -30
timeout 300 python tools/cars_spc_sweep.py 24,40,64,80,120 > gpurun_out/cars_spc.log 2>&1
cut -c1-330 gpurun_out/cars_spc.log
