#!/bin/bash
# Timing-experiment builds of libecb200 (results wrong by construction; see ECB_STRIP_EXPERIMENT in ecb_strip.cuh)
set -e
cd "$(dirname "$0")/.."; mkdir -p tools/_build
F="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared"
nvcc $F -DECB_STRIP_EXPERIMENT=1 alntools_b200/csrc/ecb_api.cu -o tools/_build/libecb_strip_noinsert.so &
nvcc $F -DECB_STRIP_EXPERIMENT=2 alntools_b200/csrc/ecb_api.cu -o tools/_build/libecb_strip_walkonly.so &
nvcc $F -DECB_MUM_PTX=1 alntools_b200/csrc/ecb_api.cu -o tools/_build/libecb_mumptx.so &
nvcc $F -DECB_WARP_PROBE=1 alntools_b200/csrc/ecb_api.cu -o tools/_build/libecb_warpprobe.so &
nvcc $F -DECB_WARP_PROBE=1 -DECB_MUM_PTX=1 alntools_b200/csrc/ecb_api.cu -o tools/_build/libecb_lean.so &
nvcc $F -DECB_WARP_PROBE=1 -DECB_MUM_PTX=1 -DECB_SMEM_BASE_ASM=1 alntools_b200/csrc/ecb_api.cu -o tools/_build/libecb_lean2.so &
nvcc $F -DECB_WARP_PROBE=1 -DECB_MUM_PTX=1 -DECB_SMEM_BASE_ASM=1 -DECB_KEY127=1 alntools_b200/csrc/ecb_api.cu -o tools/_build/libecb_lean3.so &
wait
ls -la tools/_build/
