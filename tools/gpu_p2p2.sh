mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/p2p_test.py > gpurun_out/p2p_test.log 2>&1
echo "p2p_test rc=$?" >> gpurun_out/p2p_test.log
tail -6 gpurun_out/p2p_test.log
bash tools/gpu_n8.sh 2
