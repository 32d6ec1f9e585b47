# compute-sanitizer memcheck over the tests of this round's new kernels (training, MNSRF / decoders, persistent GEMM users)
mkdir -p gpurun_out
timeout 2400 compute-sanitizer --tool memcheck --report-api-errors no --error-exitcode 3 --print-limit 20 python -m pytest tests/test_train_gpu.py tests/test_session_rankers.py -q -x -m gpu -k "not cfg2_shape" > gpurun_out/sanitize_r02_train_session.log 2>&1
echo "rc=$?" >> gpurun_out/sanitize_r02_train_session.log
timeout 2400 compute-sanitizer --tool memcheck --report-api-errors no --error-exitcode 3 --print-limit 20 python -m pytest tests/test_parity_gpu.py -q -x -m gpu -k "cars_golden or duet_golden or dssm or arc or drmm_golden or match_tensor_golden" > gpurun_out/sanitize_r02_gemm_users.log 2>&1
echo "rc=$?" >> gpurun_out/sanitize_r02_gemm_users.log
tail -4 gpurun_out/sanitize_r02_train_session.log; tail -4 gpurun_out/sanitize_r02_gemm_users.log
