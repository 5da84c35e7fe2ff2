#!/bin/bash
# builds tools/chol_bench.cpp against the built libasgfem_cuda.so (host functions only, no GPU needed) into /tmp and runs it
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
CS=$ROOT/extendableasgfem.jl_b200/csrc
OUT=${TMPDIR:-/tmp}/asgfem_chol_bench
LIBDIR=$ROOT/extendableasgfem.jl_b200
g++ -O2 -std=c++17 -pthread -I$CS -I$ROOT/include -I/usr/local/cuda/include $ROOT/tools/chol_bench.cpp -L$LIBDIR -l:libasgfem_cuda.so -Wl,-rpath,$LIBDIR -o $OUT
ASGFEM_CHOL_VERBOSE=${ASGFEM_CHOL_VERBOSE-1}
if [ -z "$ASGFEM_CHOL_VERBOSE" ]; then unset ASGFEM_CHOL_VERBOSE; else export ASGFEM_CHOL_VERBOSE; fi
$OUT "$@"
