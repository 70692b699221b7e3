#!/bin/bash
# usage: profiles/build_variant.sh NAME "-DFLAG1 -DFLAG2 ..."   ->  esdg_cns_b200/csrc/build_v/libesdg_b200_NAME.so
# Kernel-variant builds for timing experiments (only hex_fast.cu is recompiled; the other objects come from build/).
set -e
cd "$(dirname "$0")/../esdg_cns_b200/csrc"
name=$1; shift
mkdir -p build_v ../../scratch/v
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC --expt-relaxed-constexpr "$@" \
    -Xptxas -v -c -o build_v/hex_fast_$name.o hex_fast.cu 2> build_v/hex_fast_$name.ptxas.log || { tail -20 build_v/hex_fast_$name.ptxas.log; exit 1; }
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../scratch/v/libesdg_b200_$name.so build/esdg_ctx.o build_v/hex_fast_$name.o build/quad_fast.o build/peak.o build/cns1d.o -lcudart
grep -A1 "k_hex_stage_lineILi5ELb1" build_v/hex_fast_$name.ptxas.log | grep -E "registers|spill" | head -3
