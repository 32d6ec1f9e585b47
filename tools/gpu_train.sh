mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_gpu.py -x -q 2>&1 | tail -40 > gpurun_out/pytest_train.log
cat gpurun_out/pytest_train.log
