#!/bin/bash
# Evidence of the final code: all GPU tests, smoke(), bench line + reference arm, launch list, raster captures, per-CTA timeline.
set -u
mkdir -p gpurun_out
T=${1:-r2h}
timeout 1200 python -m pytest tests -m gpu -q --maxfail=12 --tb=short -p no:cacheprovider > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/${T}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
echo "bench rc=$?"; tail -c 300 gpurun_out/${T}_bench.err
timeout 400 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${T}_bench_ref.json 2>> gpurun_out/${T}_bench.err
echo "ref rc=$?"
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-network --no-train"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_bench.csv $B > gpurun_out/${T}_bench_under_ncu.log 2>&1
echo "launch list rc=$? lines=$(wc -l < gpurun_out/${T}_launches_bench.csv)"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"raster_(bin|tile)_kernel" -s 24 -c 2 -f -o gpurun_out/${T}_raster $B > gpurun_out/${T}_ncu_raster.log 2>&1
echo "raster capture rc=$?"
cp artiboost_b200/libartiboost_b200.so /tmp/lib_orig.so
cp artiboost_b200/build/variants/trace.so artiboost_b200/libartiboost_b200.so
SAMPLE_SEED=4 TAG=${T}_trace_seed4 timeout 200 python tools/trace_raster.py > gpurun_out/${T}_tile_cta_timeline_seed4.txt 2>&1
SAMPLE_SEED=1 TAG=${T}_trace_seed1 timeout 200 python tools/trace_raster.py > gpurun_out/${T}_tile_cta_timeline_seed1.txt 2>&1
cp /tmp/lib_orig.so artiboost_b200/libartiboost_b200.so
head -8 gpurun_out/${T}_tile_cta_timeline_seed4.txt; tail -2 gpurun_out/${T}_tile_cta_timeline_seed4.txt
python - <<PY
import json
d=json.load(open("gpurun_out/${T}_bench.json"))
ex=d.pop("extras")
print(json.dumps({k:d[k] for k in ("value","ms_per_step","e2e","gpu_launches","clocks","wall_s","oracle_checked","step_ms_distribution")}))
print(json.dumps({k:v for k,v in d["roofline"].items() if k!="note"}))
print({k:v for k,v in d.items() if k.startswith("train") or k.startswith("synth") or k.startswith("mpcpe")})
print(open("gpurun_out/${T}_bench_ref.json").read()[:300])
PY
