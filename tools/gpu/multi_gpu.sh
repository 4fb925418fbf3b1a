#!/bin/bash
# N-GPU: NCCL test through libkhmm's communicator + default bench at N (with the multi-GPU parity check)
cd "$(dirname "$0")/../.."
O=gpurun_out
N=${1:-2}
mkdir -p $O
nvidia-smi -L > $O/r2_g${N}_smi.txt
nvidia-smi topo -m > $O/r2_g${N}_topo.txt 2>&1
if [ "$N" = "2" ]; then
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -x -q > $O/r2_pytest_dist_2gpu.log 2>&1; echo "pytest rc=$?" >> $O/r2_pytest_dist_2gpu.log
tail -5 $O/r2_pytest_dist_2gpu.log
fi
NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 > $O/r2_bench_g$N.log 2> $O/r2_bench_g$N.err; echo "bench rc=$?" >> $O/r2_bench_g$N.err
tail -5 $O/r2_bench_g$N.err; python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_g$N.log').read().strip().splitlines()[-1])
print('N', d['n_gpus'], 'c4 ms', d['ms_per_step'], 'value %.3e'%d['value'], 'e2e ms', d['e2e']['ms_per_step'], 'h2d alone', d['e2e']['h2d_copy_alone_ms'], d['e2e'].get('host_cpu_binding'), d.get('multi_gpu_check'), d['config'].get('nccl_comm_ranks'))
for k,v in d['also'].items(): print(k, v['ms_per_step'], '%.3e'%v['value'], 'e2e', v['e2e']['ms_per_step'], v['e2e']['h2d_copy_alone_ms'])
PY
