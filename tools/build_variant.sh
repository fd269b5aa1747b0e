#!/bin/bash
# Experiment builds of the rasteriser: tools/build_variant.sh NAME "-DAB_TILE_CTAS=5 ..." -> artiboost_b200/build/variants/NAME.so
# (raster.cu recompiled with the flags, every other object from the normal build).  tools/gpu_exp.sh copies one in place per run.
set -eu
cd "$(dirname "$0")/.."
python -m artiboost_b200.build > /dev/null
mkdir -p artiboost_b200/build/variants
O=artiboost_b200/build/variants/raster_$1.o
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
  --expt-relaxed-constexpr -fmad=false -Xptxas=-v $2 -c artiboost_b200/csrc/raster.cu -o $O 2>&1 | grep -A2 "tile_kernelILi4ELb1" | grep -E "spill|Used" || true
OBJS=$(ls artiboost_b200/build/*.o | grep -v "/raster.o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o artiboost_b200/build/variants/$1.so $O $OBJS -lcuda
echo "built $1"
