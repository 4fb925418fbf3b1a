#!/bin/bash
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
timeout 900 python tools/dbg_bw3.py ${1:-131072} > $O/r2_bw3.log 2>&1; echo "rc=$?" >> $O/r2_bw3.log
tail -120 $O/r2_bw3.log
