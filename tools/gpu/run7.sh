#!/bin/bash
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
: > $O/r2_fwd_sizes.log
for b in 18944 37888 56832 131072; do
  timeout 300 python tools/dbg_bw2_time.py $b 2>&1 | grep "version 1" >> $O/r2_fwd_sizes.log
done
cat $O/r2_fwd_sizes.log
