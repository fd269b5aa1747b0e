#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 200 ncu --set full --clock-control none --import-source on -k regex:chamfer_nn_grouped -s 3 -c 1 -f -o gpurun_out/r1z_grouped_b512 python tools/prof_refiner.py > /dev/null 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:linear_f32 -s 21 -c 3 -f -o gpurun_out/r1z_linear_b512 python tools/prof_refiner.py > /dev/null 2>&1
ls -la gpurun_out | grep b512
