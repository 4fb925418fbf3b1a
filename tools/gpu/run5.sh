#!/bin/bash
# full GPU test suite + default bench + reference arm
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r2_pytest2.log 2>&1; echo "pytest rc=$?" >> $O/r2_pytest2.log
tail -5 $O/r2_pytest2.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r2_bench2.log 2> $O/r2_bench2.err; echo "bench rc=$?" >> $O/r2_bench2.err
tail -3 $O/r2_bench2.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_smoke.log 2>&1; echo "smoke rc=$?" >> $O/r2_smoke.log; tail -2 $O/r2_smoke.log
