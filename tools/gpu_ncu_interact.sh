# ncu --set full capture (with source) of the interaction kernel during a short bench run
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mt_tc_interact -s 3 -c 1 -o gpurun_out/prof_interact -f \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_interact.log 2>&1
tail -3 gpurun_out/ncu_interact.log | cut -c1-300
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench.log 2>&1
tail -1 gpurun_out/bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['stages_ms'])"
ls -la gpurun_out/
