#!/bin/bash
# Development: build kerehmm_b200/libkhmm_<name>.so with extra -D flags for ONE source file (A/B experiments on one GPU call).
# usage: tools/build_variant.sh <name> <source stem, e.g. baumwelch_tc | recursion_tc> [-DKHMM_X=1 ...]
set -e
name=$1; src=$2; shift 2
cd "$(dirname "$0")/.."
obj=kerehmm_b200/csrc/_obj
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC "$@" -c kerehmm_b200/csrc/$src.cu -o $obj/${src}_$name.o
others=$(ls $obj/*.o | grep -v "/${src}\(_[A-Za-z0-9]*\)\?\.o$" | grep -v "_[a-z0-9]*\.o$" || true)
others=$(for f in $obj/*.o; do b=$(basename $f .o); case "$b" in api|baumwelch_tc|elementwise|estep|gauss_train|recursion_generic|recursion_tc|simulate|smalln|viterbi) [ "$b" != "$src" ] && echo $f;; esac; done)
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o kerehmm_b200/libkhmm_$name.so $others $obj/${src}_$name.o
echo built kerehmm_b200/libkhmm_$name.so
