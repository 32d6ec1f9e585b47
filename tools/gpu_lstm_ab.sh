#!/bin/bash
# parity of the recurrence kernels + role timeline + headline bench after a change to lstm_tc.cu
mkdir -p gpurun_out
python -m pytest tests/test_parity_gpu.py -q -x 2>&1 | tail -4
python tools/lstm_timing.py > gpurun_out/lstm_timeline_new.txt 2>&1; head -28 gpurun_out/lstm_timeline_new.txt
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_lstm_ab.json 2> gpurun_out/bench_lstm_ab.err; python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/bench_lstm_ab.json').read().strip().splitlines()[-1])
    print({k: d[k] for k in ('value', 'ms_per_step')}, d['e2e']['value'], d['roofline'])
    print(d.get('stages'))
    for o in d.get('other_configs', []):
        print(o.get('workload', o.get('config')), o.get('value'))
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/bench_lstm_ab.err').read()[-2000:])
PY
