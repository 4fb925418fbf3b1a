#!/bin/bash
# session-2 baseline: full GPU tests, default bench (c4 + sub-records), reference arm
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/r2_smi.txt
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r2_pytest9.log 2>&1; echo "pytest rc=$?" >> $O/r2_pytest9.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r2_bench9.log 2> $O/r2_bench9.err; echo "bench rc=$?" >> $O/r2_bench9.err
tail -3 $O/r2_pytest9.log; tail -5 $O/r2_bench9.err; cat $O/r2_bench9.log
