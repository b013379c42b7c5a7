#!/bin/sh
# profiling build of the library: tools/build_prof.sh [1|2] (per-environment cycle counters in hdr[WRSN_H_PROF*]); used by tools/prof_step.py only
cd "$(dirname "$0")/.." && /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=false -std=c++17 \
  -Xcompiler -fPIC -shared -DWRSN_PROF=${1:-1} -I include -o multi_agent_rl_wrsn_b200/csrc/libwrsn_b200_prof${1:-1}.so multi_agent_rl_wrsn_b200/csrc/wrsn_kernels.cu
