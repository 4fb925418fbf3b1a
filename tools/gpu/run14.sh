#!/bin/bash
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
QUICK=1 timeout 900 python tools/dbg_fw3.py ${1:-37888} > $O/r2_fw3.log 2>&1; echo "rc=$?" >> $O/r2_fw3.log; grep "stash\|FAIL\|Error\|error" $O/r2_fw3.log | head -20; tail -9 $O/r2_fw3.log | cut -c1-1000
