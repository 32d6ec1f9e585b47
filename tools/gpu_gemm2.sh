mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "duet or dssm or arc or cars or gemm" 2>&1 | tail -12 > gpurun_out/pytest_gemm2.log
tail -6 gpurun_out/pytest_gemm2.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_gemm2.log 2>&1
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_gemm2.log').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'])
for o in d.get('other_configs', []):
    print(' ', o.get('name'), round(o.get('pairs_per_s', 0)), o.get('stages_ms'))
PY
