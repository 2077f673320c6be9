#!/bin/bash
# usage: build_variant.sh NK RING SKEW -> libnerf_b200_nk${NK}_r${RING}_s${SKEW}.so (development sweeps)
set -e
cd "$(dirname "$0")"
NK=$1; RING=$2; SKEW=$3; shift 3
OUT=variants/libnerf_b200_nk${NK}_r${RING}_s${SKEW}.so
mkdir -p variants
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr \
  -DNERF_TC_NK=$NK -DNERF_TC_RING=$RING -DNERF_TC_SKEW=$SKEW "$@" -shared -o $OUT nerf_api.cu nerf_render.cu nerf_mlp_fp32.cu nerf_mlp_tc.cu nerf_mlp_wgrad.cu nerf_train.cu nerf_data.cu
echo built $OUT
