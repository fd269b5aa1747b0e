#!/bin/bash
set -u
mkdir -p gpurun_out
T=u9
timeout 900 python -m pytest tests -m gpu -q --maxfail=12 --tb=short -p no:cacheprovider -k "network or trainstep or train_ops or conv or gemm or train" 2>&1 | tail -5 | tee gpurun_out/${T}_pytest.log
for p in 0 1; do
  AB_GEMM_PERSISTENT=$p python tools/time_gemm.py 2>&1 | tee -a gpurun_out/${T}_gemm.txt
  AB_GEMM_PERSISTENT=$p python tools/time_train_step.py 2>&1 | tail -1 | tee -a gpurun_out/${T}_step.txt
done
