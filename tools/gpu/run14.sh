#!/bin/bash
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
timeout 900 python tools/dbg_fw3.py ${1:-37888} > $O/r2_fw3.log 2>&1; echo "rc=$?" >> $O/r2_fw3.log; grep "stash\|FAIL\|Error\|error" $O/r2_fw3.log | head -20; grep "forward version" $O/r2_fw3.log
timeout 300 python tools/dbg_bw_trace.py 18944 2>&1 | grep FWD | tail -2 | cut -c1-600
