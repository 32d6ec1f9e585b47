# 2-GPU check of cair_allgather_scores (peer stores over NVLink) against NCCL, then the N=2 bench with it and with NCCL.
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/p2p_test.py > gpurun_out/p2p_test.log 2>&1
echo "p2p_test rc=$?" >> gpurun_out/p2p_test.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --no-cpu-baseline > gpurun_out/bench_n2_p2p.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --no-cpu-baseline --nccl-collective --no-other-configs > gpurun_out/bench_n2_nccl.log 2>&1
tail -8 gpurun_out/p2p_test.log; tail -1 gpurun_out/bench_n2_p2p.log | cut -c1-3000; tail -1 gpurun_out/bench_n2_nccl.log | cut -c1-1200
