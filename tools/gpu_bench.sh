mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -3 gpurun_out/bench_n1.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n1.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","gpu_launches","stages_ms","clocks")})
print(d["e2e"]); print(d["roofline"]); print(d.get("cpu_baseline"))
for o in d["other_configs"]: print(o["name"], round(o["pairs_per_s"]), round(o["ms_per_step"],3), o["roofline"]["bound"], round(o["roofline"]["frac"],4), o["stages_ms"])
PY
