#!/bin/bash
# builds tools/chol_bench.cpp against csrc/chol.cpp + dense_chol.cpp (host only) into /tmp and runs it
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
CS=$ROOT/extendableasgfem.jl_b200/csrc
OUT=${TMPDIR:-/tmp}/asgfem_chol_bench
g++ -O3 -std=c++17 -pthread -I$CS -I$ROOT/include -I/usr/local/cuda/include $ROOT/tools/chol_bench.cpp $CS/chol.cpp $CS/dense_chol.cpp -o $OUT
ASGFEM_CHOL_VERBOSE=${ASGFEM_CHOL_VERBOSE-1} $OUT "$@"
