#!/bin/bash
# Builds the warp-specialised K2p experiment: build_k2p_ws.sh <suffix> [-DMCBA_WS_FENCE=0 -DMCBA_WS_STAGES=8 ...]  ->  k2p_ws_<suffix>
set -e
cd "$(dirname "$0")"
sfx=$1; shift
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -I../../include -I../../multicam_calibration_b200/csrc \
  -I. -Xptxas -v "$@" -o k2p_ws_$sfx k2p_ws_bench.cu 2> k2p_ws_$sfx.ptxas.log
grep -A2 "k2p_ws_kernel" k2p_ws_$sfx.ptxas.log | grep "spill\|Used" | head -2
