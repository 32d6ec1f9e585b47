mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -5 gpurun_out/bench_n2.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n2.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","n_gpus","gpu_launches","stages_ms")})
print(d["e2e"])
for o in d["other_configs"]: print(o["name"], round(o["pairs_per_s"]), round(o["ms_per_step"],3), o["n_gpus"])
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>&1 | tail -2 | cut -c1-400
