#!/bin/bash
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
QUICK=1 timeout 900 python tools/dbg_bw3.py 131072 > $O/r2_bw3_full.log 2>&1; tail -28 $O/r2_bw3_full.log | grep -v "^step 13\|^step 14" | cut -c1-1200
