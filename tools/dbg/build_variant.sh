#!/usr/bin/env bash
# build a variant of the library with extra nvcc flags for hp2 (never shipped; loaded via AGX_LIB_PATH)
#   [HP1_FLAGS=...] tools/dbg/build_variant.sh <name> <extra flags for hp2_raycast.cu...>
set -e
cd "$(dirname "$0")/../.."
NAME=$1; shift
C=aerial_gym_simulator_b200/csrc
F="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr"
B=tools/dbg/build_$NAME
mkdir -p $B
nvcc $F -c $C/agx_common.cu -o $B/c.o
nvcc $F -prec-div=false -prec-sqrt=false $HP1_FLAGS -Xptxas -v -c $C/hp1.cu -o $B/h1.o 2>&1 | grep -A3 "step_kernelILi4ELb1ELb1" | grep "Used\|spill"
nvcc $F -c $C/hp1_aux.cu -o $B/h1a.o
nvcc $F -fmad=false "$@" -Xptxas -v -c $C/hp2_raycast.cu -o $B/h2.o 2>&1 | grep -A3 "cast_kernelILb1ELb1" | grep "Used\|spill"
nvcc $F -c $C/p2p_allgather.cu -o $B/p.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o tools/dbg/libagx_$NAME.so $B/*.o
echo built tools/dbg/libagx_$NAME.so
