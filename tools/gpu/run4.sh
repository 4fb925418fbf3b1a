#!/bin/bash
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
timeout 900 python tools/dbg_split.py 131072 > $O/r2_dbg_split.log 2>&1; echo "rc=$?" >> $O/r2_dbg_split.log
tail -20 $O/r2_dbg_split.log
