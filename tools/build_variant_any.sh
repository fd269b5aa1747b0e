#!/bin/bash
# Experiment build of one source file: tools/build_variant_any.sh NAME FILE.cu "-DFLAG=.." -> artiboost_b200/build/variants/NAME.so
set -eu
cd "$(dirname "$0")/.."
python -m artiboost_b200.build > /dev/null
mkdir -p artiboost_b200/build/variants
F=$2; B=$(basename $F .cu)
O=artiboost_b200/build/variants/${B}_$1.o
EXTRA=""; case $B in raster|augment) EXTRA="-fmad=false";; esac
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
  --expt-relaxed-constexpr $EXTRA -Xptxas=-v $3 -c artiboost_b200/csrc/$B.cu -o $O 2>&1 | grep -E "spill|Used" | sort | uniq -c | sort -rn | head -4 || true
OBJS=$(ls artiboost_b200/build/*.o | grep -v "/$B.o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o artiboost_b200/build/variants/$1.so $O $OBJS -lcuda
echo "built $1"
