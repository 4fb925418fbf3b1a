#!/bin/bash
# ncu full-set capture of ONE launch of the small-N log-likelihood kernel at config-2 shape (after the program ran clean).
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
timeout 300 python tools/loglik_c2_time.py > $O/r2_loglik_c2_time.log 2>&1 || exit 1
cat $O/r2_loglik_c2_time.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fwdbwd_ --launch-skip 5 -c 1 -f -o $O/r2_ncu_c2_fwdbwd python tools/loglik_c2_time.py > $O/r2_ncu_c2.log 2>&1
ls -la $O/r2_ncu_c2_fwdbwd.ncu-rep
