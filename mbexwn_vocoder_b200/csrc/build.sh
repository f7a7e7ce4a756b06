#!/bin/bash
# Build libmbexwn_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU).
set -e
cd "$(dirname "$0")"
OUT=../libmbexwn_b200.so
nvcc --threads 0 -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
     -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -shared -cudart static \
     ${MBX_PTXAS_V:+-Xptxas -v} \
     -o "$OUT" api.cu k_conv.cu k_excitation.cu k_synth.cu k_norm.cu k_wavenet_tc.cu k_wavenet_layer.cu
echo "built $(realpath $OUT)"
