#!/bin/bash
set -u
mkdir -p gpurun_out
python tools/trace_synth_pipeline.py 2>&1 | tee gpurun_out/u2_trace.txt
