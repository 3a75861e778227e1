#!/bin/bash
# One GPU-box visit for the pair-body micro-benchmark (tools/pair_body_bench.cu): which issue order of the per-pair
# arithmetic of k_demux_default's dictionary variant is fastest at 8 and 12 warps per SM.  Usage (under gpurun):
#   bash tools/gpu_body.sh <tag>
TAG=${1:-body}
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/pair_body_bench tools/pair_body_bench.cu || exit 1
timeout 120 /tmp/pair_body_bench | tee gpurun_out/${TAG}_pair_body_bench.txt
