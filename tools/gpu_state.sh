mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 900 python bench.py > gpurun_out/bench.log 2>&1
tail -5 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print('value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], d.get('stages_ms'))
print('cpu_baseline', d.get('cpu_baseline'))
for o in d.get('other_configs', []):
    print(' ', o.get('name'), round(o.get('pairs_per_s', 0)), o.get('roofline', {}).get('kernel'), round(o.get('roofline', {}).get('frac', 0), 4), o.get('stages_ms'))
PY
