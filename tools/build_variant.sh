#!/bin/bash
# Development: build kerehmm_b200/libkhmm_<name>.so with extra -D flags for ONE source file (A/B experiments on one GPU call).
# usage: tools/build_variant.sh <name> <source stem, e.g. baumwelch_tc | recursion_tc> [-DKHMM_X=1 ...]
set -e
name=$1; src=$2; shift 2
cd "$(dirname "$0")/.."
obj=kerehmm_b200/csrc/_obj
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC "$@" -c kerehmm_b200/csrc/$src.cu -o $obj/${src}__$name.o
others=$(for f in kerehmm_b200/csrc/*.cu; do b=$(basename $f .cu); [ "$b" != "$src" ] && echo $obj/$b.o; done)
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o kerehmm_b200/libkhmm_$name.so $others $obj/${src}__$name.o -ldl
echo built kerehmm_b200/libkhmm_$name.so
