#!/bin/bash
# Development: build kerehmm_b200/libkhmm_<name>.so with extra -D flags for baumwelch_tc.cu (A/B experiments on one GPU call).
# usage: tools/build_variant.sh <name> [-DKHMM_X=1 ...]
set -e
name=$1; shift
cd "$(dirname "$0")/.."
obj=kerehmm_b200/csrc/_obj
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC "$@" -c kerehmm_b200/csrc/baumwelch_tc.cu -o $obj/baumwelch_tc_$name.o
others=$(ls $obj/*.o | grep -v baumwelch_tc)
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o kerehmm_b200/libkhmm_$name.so $others $obj/baumwelch_tc_$name.o
echo built kerehmm_b200/libkhmm_$name.so
