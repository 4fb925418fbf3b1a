#!/bin/bash
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
timeout 900 python tools/dbg_bw3.py ${1:-37888} > $O/r2_bw3.log 2>&1; grep -v "nseq\|flags\|loglik" $O/r2_bw3.log | grep "v3 vs v1" | awk '{print $1,$2,$6}' | sort -k3 -g | tail -6; tail -28 $O/r2_bw3.log | grep -v "full size" | cut -c1-1400
