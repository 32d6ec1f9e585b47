mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests -m gpu -q -x -k "lstm_entry_point_vs_oracle and cluster and 37-23-40-64 and LSTM" > gpurun_out/sanitize_rnn.log 2>&1
grep -v "^$" gpurun_out/sanitize_rnn.log | head -60
