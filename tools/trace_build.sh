#!/bin/bash
# Build the TRACE variant of the library (per-role wait counters compiled in: -DSTCD_TRACE_WAITS) next to the product one.
# usage: bash tools/trace_build.sh && STCD_LIB=stcd_b200/libstcd_b200_trace.so python tools/trace_op.py 64
cd "$(dirname "$0")/.."
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared -cudart static --threads 0 -DSTCD_TRACE_WAITS \
  -o stcd_b200/libstcd_b200_trace.so stcd_b200/csrc/*.cu
