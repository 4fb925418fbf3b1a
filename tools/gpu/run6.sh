#!/bin/bash
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_train_loop.py tests/test_gpu_parity.py -m gpu -x -q -k "tensor_core or config4 or transposed or device_resident" > $O/r2_pytest3.log 2>&1; echo "pytest rc=$?" >> $O/r2_pytest3.log
tail -4 $O/r2_pytest3.log
timeout 900 python tools/dbg_split.py 131072 > $O/r2_dbg_split.log 2>&1; echo "rc=$?" >> $O/r2_dbg_split.log
tail -9 $O/r2_dbg_split.log
timeout 300 python tools/dbg_bw_trace.py 18944 > $O/r2_bw_trace1.log 2>&1; tail -8 $O/r2_bw_trace1.log
