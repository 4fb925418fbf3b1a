#!/bin/bash
# full GPU tests, default bench (c4 + sub-records)
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > $O/r2_smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2_pytest9.log 2>&1; echo "pytest rc=$?" >> $O/r2_pytest9.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r2_bench9.log 2> $O/r2_bench9.err; echo "bench rc=$?" >> $O/r2_bench9.err
tail -5 $O/r2_pytest9.log; tail -5 $O/r2_bench9.err; cat $O/r2_smi.txt; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench9.log').read().strip().splitlines()[-1])
r=d['roofline']
print('c4 ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], 'bwd', r['kernel_ms'], 'fwd', r['forward_kernel']['kernel_ms'], 'frac', r['frac'], 'other', r['operand_precision']['other'], r['operand_precision']['other_ms_per_step'], d['clocks'])
for k,v in d['also'].items(): print(k, v['ms_per_step'], v['roofline']['frac'])
PY
