#!/bin/bash
# ncu --set full capture (with source) of one raster kernel of the bench command.  usage: gpu_ncu.sh TAG KERNEL_REGEX
set -u
mkdir -p gpurun_out
T=$1; K=$2
timeout 300 ncu --set full --clock-control none --cache-control none --import-source on -k regex:$K -s 24 -c 1 -f -o gpurun_out/${T}_$K python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-network --no-train > /dev/null 2>&1
ls -la gpurun_out/${T}_$K.ncu-rep
