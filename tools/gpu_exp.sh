#!/bin/bash
# Raster experiments: device timing of the bench workload under env switches and experiment builds (tools/build_variant.sh).
# usage: gpu_exp.sh TAG "VARIANT=name ENV1=..;VARIANT=.. ENV2=.." [parity]  -- with `parity`, the raster parity tests run per variant
set -u
mkdir -p gpurun_out
T=${1:-x}
IFS=';' read -ra VARS <<< "${2:-VARIANT=base}"
cp artiboost_b200/libartiboost_b200.so /tmp/lib_orig.so
last=""
for v in "${VARS[@]}"; do
  var=$(echo "$v" | tr ' ' '\n' | grep '^VARIANT=' | cut -d= -f2)
  if [ -n "$var" ]; then cp artiboost_b200/build/variants/$var.so artiboost_b200/libartiboost_b200.so; else cp /tmp/lib_orig.so artiboost_b200/libartiboost_b200.so; fi
  echo "== $v" | tee -a gpurun_out/${T}_time.log
  if [ "${3:-}" = "parity" ] && [ "$var" != "$last" ]; then
    env $v timeout 600 python -m pytest tests -m gpu -q --maxfail=5 --tb=short -p no:cacheprovider -k "raster" 2>&1 | tail -3
    last=$var
  fi
  env $v N=300 timeout 200 python tools/time_raster.py 2>&1 | grep -E "graph|us per launch" | tee -a gpurun_out/${T}_time.log
done
cp /tmp/lib_orig.so artiboost_b200/libartiboost_b200.so
