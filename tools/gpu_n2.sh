mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --no-cpu-baseline > gpurun_out/bench_n2.log 2>&1
echo rc=$?
python - <<'PY'
import json
lines=[l for l in open('gpurun_out/bench_n2.log').read().strip().splitlines() if l.startswith('{')]
if not lines:
    print(open('gpurun_out/bench_n2.log').read()[-3000:])
else:
    d = json.loads(lines[-1])
    print('value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], d['run'].get('collective'))
    for o in d.get('other_configs', []):
        print(' ', o.get('name'), round(o.get('pairs_per_s', 0)), o.get('n_gpus'))
PY
timeout 300 python -m pytest tests/test_p2p_gpu.py -q 2>&1 | tail -2
