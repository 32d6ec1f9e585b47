mkdir -p gpurun_out
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --no-cpu-baseline > gpurun_out/bench_n$N.log 2>&1
echo rc=$?
python - $N <<'PY'
import json, sys
n = sys.argv[1]
lines=[l for l in open('gpurun_out/bench_n%s.log' % n).read().strip().splitlines() if l.startswith('{')]
if not lines:
    print(open('gpurun_out/bench_n%s.log' % n).read()[-3000:])
else:
    d = json.loads(lines[-1])
    print('value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], d['run'].get('collective'), d['run'].get('collective_fallback_reason'))
    for o in d.get('other_configs', []):
        print(' ', o.get('name'), round(o.get('pairs_per_s', 0)), o.get('n_gpus'))
PY
