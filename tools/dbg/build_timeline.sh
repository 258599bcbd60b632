#!/usr/bin/env bash
# debug variant of the library with per-warp timestamps (never shipped; loaded via AGX_LIB_PATH)
set -e
cd "$(dirname "$0")/../.."
C=aerial_gym_simulator_b200/csrc
F="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr"
mkdir -p tools/dbg/build
nvcc $F -c $C/agx_common.cu -o tools/dbg/build/c.o
nvcc $F -prec-div=false -prec-sqrt=false -DAGX_TIMELINE -c $C/hp1.cu -o tools/dbg/build/h1.o
nvcc $F -fmad=false -c $C/hp2_raycast.cu -o tools/dbg/build/h2.o
nvcc $F -c $C/p2p_allgather.cu -o tools/dbg/build/p.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o tools/dbg/libagx_timeline.so tools/dbg/build/*.o
echo built tools/dbg/libagx_timeline.so
