#!/bin/bash
# Regenerates the round-2 evidence under gpurun_out/ (copied into profiles/ afterwards): ncu launch list of the bench command,
# full captures of the raster kernels and of the new training-step kernels, kernel list of one training step, per-CTA
# timeline of the tile kernel (needs artiboost_b200/build/variants/trace.so: tools/build_variant.sh trace -DAB_RASTER_TRACE).
set -u
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-network --no-train"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv $B > gpurun_out/r2_bench_under_ncu.log 2>&1
echo "launch list rc=$? lines=$(wc -l < gpurun_out/r2_launches_bench.csv)"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"raster_(bin|tile)_kernel" -s 8 -c 2 -f -o gpurun_out/r2_raster $B > gpurun_out/r2_ncu_raster.log 2>&1
echo "raster capture rc=$?"
N=2 timeout 500 ncu --set full --clock-control none --import-source on -k regex:"tail_loss_kernel|head_decode_online_kernel|head_decode_bwd_lse_kernel|bn_bwd_reduce_kernel" -c 6 -f -o gpurun_out/r2_train_new python tools/time_train_step.py > gpurun_out/r2_ncu_train.log 2>&1
echo "train capture rc=$?"
timeout 300 python tools/prof_step_kernels.py > gpurun_out/r2_train_step_kernels.txt 2>&1
echo "step kernels rc=$?"; head -4 gpurun_out/r2_train_step_kernels.txt | tail -2
if [ -f artiboost_b200/build/variants/trace.so ]; then
  cp artiboost_b200/libartiboost_b200.so /tmp/lib_orig.so
  cp artiboost_b200/build/variants/trace.so artiboost_b200/libartiboost_b200.so
  TAG=r2_tile_trace timeout 200 python tools/trace_raster.py > gpurun_out/r2_tile_cta_timeline.txt 2>&1
  echo "trace rc=$?"; head -12 gpurun_out/r2_tile_cta_timeline.txt
  cp /tmp/lib_orig.so artiboost_b200/libartiboost_b200.so
fi
