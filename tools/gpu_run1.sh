#!/bin/bash
# One gpurun call: GPU parity tests, the bench line, the ncu launch list of the same bench command and full captures of
# the dominant kernels.  Outputs under gpurun_out/ (scratch; summaries are copied to profiles/ by hand).
set -u
mkdir -p gpurun_out
T=${1:-r1c}
timeout 900 python -m pytest tests -m gpu -q --maxfail=12 --tb=short -p no:cacheprovider > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/${T}_pytest.log
tail -5 gpurun_out/${T}_pytest.log
timeout 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
echo "bench rc=$?"; tail -c 600 gpurun_out/${T}_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_ref.json 2>> gpurun_out/${T}_bench.err
NCU_BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-network --no-train"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/${T}_launches_raster.csv $NCU_BENCH > /dev/null 2>&1
for k in raster_triangle raster_resolve; do
  timeout 300 ncu --set full --clock-control none --cache-control none --import-source on -k regex:$k -s 24 -c 2 -f -o gpurun_out/${T}_$k $NCU_BENCH > /dev/null 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:chamfer_nn -c 2 -f -o gpurun_out/${T}_chamfer_nn python -m pytest tests/test_gpu_refine.py -q -k full_batch -p no:cacheprovider > /dev/null 2>&1
ls -la gpurun_out | tail -12
timeout 300 ncu --set full --clock-control none --import-source on -k regex:linear_f32 -s 3 -c 2 -f -o gpurun_out/${T}_linear_f32 python -m pytest tests/test_gpu_refine.py -q -k "synth_pipeline and hand_obj" -p no:cacheprovider > /dev/null 2>&1
ls -la gpurun_out | tail -5
