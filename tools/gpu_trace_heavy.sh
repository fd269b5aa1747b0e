#!/bin/bash
cp artiboost_b200/libartiboost_b200.so /tmp/lib_orig.so
cp artiboost_b200/build/variants/trace.so artiboost_b200/libartiboost_b200.so
SAMPLE_SEED=4 TAG=heavy_trace timeout 200 python tools/trace_raster.py 2>&1 | tee gpurun_out/heavy_trace.txt | tail -32
cp /tmp/lib_orig.so artiboost_b200/libartiboost_b200.so
