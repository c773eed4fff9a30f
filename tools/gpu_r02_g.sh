#!/bin/bash
# 2-GPU call: NCCL gradient-equality test + the N=2 bench line (train block + ddp_check)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
nvidia-smi -L > gpurun_out/r02g_gpus.txt
timeout 600 python -m pytest -q -p no:cacheprovider tests/test_dist_gpu.py -m gpu -s > gpurun_out/r02g_dist_gpu_n2.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02g_dist_gpu_n2.log
tail -5 gpurun_out/r02g_dist_gpu_n2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 10 --warmup 3 \
  > gpurun_out/r02g_bench_n2.json 2> gpurun_out/r02g_bench_n2.err
tail -c 1500 gpurun_out/r02g_bench_n2.json; tail -5 gpurun_out/r02g_bench_n2.err
