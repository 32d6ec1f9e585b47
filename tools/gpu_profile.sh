# usage: bash tools/gpu_profile.sh <kernel-regex> <tag>   -> ncu --set full capture of one kernel of a short bench run
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$1 -s 2 -c 1 -o gpurun_out/prof_$2 -f \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$2.log 2>&1
tail -3 gpurun_out/ncu_$2.log | cut -c1-400
ls -la gpurun_out/
