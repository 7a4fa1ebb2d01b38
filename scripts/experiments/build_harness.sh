#!/bin/bash
# Build the clock64 / timing harness (scripts/trace_assign.cu) for HEAD and for every patch in this directory, into
# scripts/_bin/ (git-ignored, but shipped to the GPU box by gpurun).  Run here (nvcc cross-compiles), then e.g.
#   gpurun --timeout 300 -- 'for b in head fused_gather_warps_setmaxnreg three_epilogue_groups_setmaxnreg; do
#       echo == $b; timeout 60 scripts/_bin/trace_assign_$b 16 0 | head -3; timeout 60 scripts/_bin/trace_assign_$b 16 1 | head -3; done'
# (always wrap GPU commands in your own `timeout`: a hung 8-GPU run spent the whole round-1 budget).
set -e
cd "$(dirname "$0")/../.."
NVCC="nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -lcuda"
mkdir -p scripts/_bin
$NVCC -o scripts/_bin/trace_assign_head scripts/trace_assign.cu
for p in scripts/experiments/*.patch; do
  name=$(basename "$p" .patch)
  git apply "$p"
  $NVCC -o "scripts/_bin/trace_assign_$name" scripts/trace_assign.cu || true
  git apply -R "$p"
done
ls -la scripts/_bin/trace_assign_*
