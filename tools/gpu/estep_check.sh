#!/bin/bash
# fused E-step: backward kernel 3 against kernel 1 and the oracle, kernel times (QUICK=1: one shape), timeline of block 0
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
timeout 900 python tools/estep_backward_check.py ${1:-37888} > $O/r2_bw3.log 2>&1; echo "rc=$?" >> $O/r2_bw3.log
grep -v "nseq\|flags\|loglik" $O/r2_bw3.log | grep "v3 vs v1" | awk '{print $1,$2,$6}' | sort -k3 -g | tail -5; grep "forward/backward\|rc=\|period" $O/r2_bw3.log
