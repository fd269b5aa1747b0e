#!/bin/bash
set -u
mkdir -p gpurun_out
python tools/trace_synth_pipeline.py 2>&1 | grep -E "no marks|host" | tee gpurun_out/u4_trace.txt
AB_SYNTH_PRIO=0 python tools/trace_synth_pipeline.py 2>&1 | grep -E "no marks|host" | tee -a gpurun_out/u4_trace.txt
python tools/time_bn.py 2>&1 | tail -2 | tee gpurun_out/u4_bn.txt
timeout 600 python -m pytest tests -m gpu -q --maxfail=12 --tb=short -p no:cacheprovider -k "network or trainstep or train_ops or conv or bn" 2>&1 | tail -8 | tee gpurun_out/u4_pytest.log
python tools/time_train_step.py 2>&1 | tail -5 | tee gpurun_out/u4_step.txt
