#!/bin/bash
# usage: profiles/build_variant_quad.sh NAME "-DFLAG ..."  ->  scratch/v/libesdg_b200_NAME.so (only quad_fast.cu is recompiled)
set -e
cd "$(dirname "$0")/../esdg_cns_b200/csrc"
name=$1; shift
mkdir -p build_v ../../scratch/v
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC --expt-relaxed-constexpr "$@" \
    -Xptxas -v -c -o build_v/quad_fast_$name.o quad_fast.cu 2> build_v/quad_fast_$name.ptxas.log || { tail -20 build_v/quad_fast_$name.ptxas.log; exit 1; }
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../scratch/v/libesdg_b200_$name.so build/esdg_ctx.o build/hex_fast.o build_v/quad_fast_$name.o build/peak.o build/cns1d.o -lcudart
grep -A2 "k_quad_stage_lineILi4" build_v/quad_fast_$name.ptxas.log | grep -E "registers|spill" | head -4
