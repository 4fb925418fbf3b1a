#!/bin/bash
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
: > $O/r2_variants.log
for v in "" $VARIANTS; do
  KHMM_VARIANT=$v timeout 300 python tools/dbg_bw2_time.py 131072 >> $O/r2_variants.log 2>&1
done
cat $O/r2_variants.log
