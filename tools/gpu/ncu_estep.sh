#!/bin/bash
# ncu evidence for config 4 (run only after the profiled programs have exited 0 without a profiler):
#   1. full-set captures of ONE launch of the backward and of the forward E-step kernel, split bf16 operands (the default),
#      at B = 18,944 (one tile per SM), with source correlation;
#   2. the launch list (gpu__time_duration) of a short default bench.py run.
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
timeout 300 python tools/estep_kernel_times.py 18944 > $O/r2_estep_times_18944.log 2>&1 || exit 1
cat $O/r2_estep_times_18944.log
# estep_kernel_times.py launches each kernel 5 times with tf32 operands, then 5 times with split bf16 operands
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bw_backward_tc3 --launch-skip 5 -c 1 -f -o $O/r2_ncu_c4_bw_backward_tc3 python tools/estep_kernel_times.py 18944 > $O/r2_ncu_bwd.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bw_forward_tc_kernel --launch-skip 5 -c 1 -f -o $O/r2_ncu_c4_bw_forward_tc python tools/estep_kernel_times.py 18944 > $O/r2_ncu_fwd.log 2>&1
timeout 600 python bench.py --steps 3 --warmup 3 --also none --no-cpu > $O/r2_bench_short.log 2>&1 || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_c4_baumwelch_tc.csv python bench.py --steps 3 --warmup 3 --also none --no-cpu > $O/r2_ncu_launches.log 2>&1
ls -la $O/*.ncu-rep $O/r2_launches_c4_baumwelch_tc.csv
