#!/bin/bash
# GPU call 1 (round 2): tests, default bench (c4 + sub-records), microbenchmarks, backward-kernel timeline
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/r2_smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2_pytest1.log 2>&1; echo "pytest rc=$?" >> $O/r2_pytest1.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r2_bench1.log 2> $O/r2_bench1.err; echo "bench rc=$?" >> $O/r2_bench1.err
timeout 120 tools/microbench_mma > $O/r2_microbench_tf32.json 2> $O/r2_microbench_mma.err
timeout 120 tools/microbench_red > $O/r2_microbench_red.txt 2>&1
timeout 300 python tools/dbg_bw_trace.py 18944 > $O/r2_bw_trace0.log 2>&1
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/r2_bench_ref.log 2>&1
tail -3 $O/r2_pytest1.log; cat $O/r2_bench1.err | tail -5; cat $O/r2_microbench_tf32.json; cat $O/r2_microbench_red.txt; tail -12 $O/r2_bw_trace0.log
