#!/bin/bash
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
QUICK=1 timeout 600 python tools/dbg_bw2.py 131072 > $O/r2_dbg_bw2.log 2>&1; echo "rc=$?" >> $O/r2_dbg_bw2.log
tail -24 $O/r2_dbg_bw2.log
