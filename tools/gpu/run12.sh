#!/bin/bash
# ncu capture of the tc3 backward kernel (one launch), full set with source correlation
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
KHMM_VER=3 timeout 300 python tools/dbg_bw3_time.py 18944 > $O/r2_bw3_pre_ncu.log 2>&1 || exit 1
KHMM_VER=3 timeout 900 ncu --set full --clock-control none --import-source on -k regex:bw_backward_tc3 -c 1 -f -o $O/r2_bw3 python tools/dbg_bw3_time.py 18944 > $O/r2_bw3_ncu.log 2>&1
tail -3 $O/r2_bw3_ncu.log; ls -la $O/r2_bw3.ncu-rep
