# compute-sanitizer memcheck over the whole GPU test suite
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 99 --print-limit 20 \
  python -m pytest tests -m gpu -q -x > gpurun_out/sanitize.log 2>&1
echo "exit $?" >> gpurun_out/sanitize.log
grep -E "ERROR SUMMARY|passed|failed|exit|Invalid|out of bounds" gpurun_out/sanitize.log | head -20
tail -5 gpurun_out/sanitize.log
