#!/bin/bash
# 2-GPU: NCCL test through libkhmm's communicator + default bench at N=2 (with the multi-GPU parity check)
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
nvidia-smi -L > $O/r2_g2_smi.txt
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -x -q > $O/r2_pytest_dist_2gpu.log 2>&1; echo "pytest rc=$?" >> $O/r2_pytest_dist_2gpu.log
tail -5 $O/r2_pytest_dist_2gpu.log
NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 > $O/r2_bench_g2.log 2> $O/r2_bench_g2.err; echo "bench rc=$?" >> $O/r2_bench_g2.err
tail -5 $O/r2_bench_g2.err; tail -c 1500 $O/r2_bench_g2.log
