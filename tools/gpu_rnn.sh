# recurrence-focused pass: new-kernel parity first (bounded), then the suite, bench A/B, role timing. logs under gpurun_out/
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x -k "lstm_entry or rnn_uni or selftest" 2>&1 | tail -25 > gpurun_out/pytest_rnn.log
cat gpurun_out/pytest_rnn.log | tail -12
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench.log 2>&1
tail -1 gpurun_out/bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['stages_ms'])"
timeout 300 python tools/rnn_timing.py > gpurun_out/rnn_timing.log 2>&1
cat gpurun_out/rnn_timing.log | tail -30
