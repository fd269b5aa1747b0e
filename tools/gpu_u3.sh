#!/bin/bash
set -u
mkdir -p gpurun_out
python tools/trace_synth_pipeline.py 2>&1 | tee gpurun_out/u2_trace.txt
for v in "AB_BN_TUNE=0 AB_IM2COL_PAIRS=0" "AB_BN_TUNE=1 AB_IM2COL_PAIRS=1"; do
  env $v python tools/time_bn.py 2>&1 | tee -a gpurun_out/u3_bn.txt
done
timeout 600 python -m pytest tests -m gpu -q --maxfail=12 --tb=short -p no:cacheprovider -k "network or trainstep or train_ops or conv or bn" 2>&1 | tail -8 | tee gpurun_out/u3_pytest.log
