#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "$1" > gpurun_out/r2_one_test.log 2>&1; echo "rc=$?" >> gpurun_out/r2_one_test.log; tail -25 gpurun_out/r2_one_test.log
