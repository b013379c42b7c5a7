#!/bin/sh
# tuning builds of the observation raster: tools/build_obs_variants.sh  -> csrc/libwrsn_b200_obs<k>.so (used by tools/time_observe.py only)
cd "$(dirname "$0")/.."
k=0
for v in "-DOBS_TJ32=20 -DOBS_THREADS32=128 -DOBS_CH32=64 -DOBS_MINB32=3" "-DOBS_TJ32=20 -DOBS_THREADS32=128 -DOBS_CH32=32 -DOBS_MINB32=3" "-DOBS_TJ32=20 -DOBS_THREADS32=128 -DOBS_CH32=32 -DOBS_MINB32=4" "-DOBS_TJ32=10 -DOBS_THREADS32=256 -DOBS_CH32=32 -DOBS_MINB32=3" "-DOBS_TJ32=10 -DOBS_THREADS32=256 -DOBS_CH32=64 -DOBS_MINB32=2" "-DOBS_TJ32=10 -DOBS_THREADS32=256 -DOBS_CH32=32 -DOBS_MINB32=4"; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=false -std=c++17 -Xcompiler -fPIC -shared $v -I include \
    -o multi_agent_rl_wrsn_b200/csrc/libwrsn_b200_obs$k.so multi_agent_rl_wrsn_b200/csrc/wrsn_kernels.cu &
  echo "$k: $v"
  k=$((k+1))
done
wait
