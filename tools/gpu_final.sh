#!/bin/bash
# Final verification of the round: all GPU tests, smoke(), the bench line + reference arm, the ncu launch list of the same
# bench command and one capture of the grouped nearest-neighbour kernel.
set -u
mkdir -p gpurun_out
T=${1:-r1z}
timeout 900 python -m pytest tests -m gpu -q --maxfail=12 --tb=short -p no:cacheprovider > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/${T}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
echo "bench rc=$?"; tail -c 400 gpurun_out/${T}_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_ref.json 2>> gpurun_out/${T}_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/${T}_launches_raster.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-network --no-train > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:chamfer_nn_grouped -s 1 -c 1 -f -o gpurun_out/${T}_chamfer_nn_grouped python -m pytest tests/test_gpu_refine.py -q -k "grouped_is_bit and 778-10000" -p no:cacheprovider > /dev/null 2>&1
python - <<PY
import json
d=json.load(open("gpurun_out/${T}_bench.json"))
ex=d.pop("extras")
print(json.dumps({k:d[k] for k in ("value","ms_per_step","e2e","gpu_launches","clocks")}))
print(json.dumps({k:v for k,v in d["roofline"].items() if k!="note"}))
r=ex.get("hand_obj_refiner_8f3"); print({k:v for k,v in r.items() if not k.startswith("stage") and k!="chamfer_nn"})
t=ex.get("train_loop_configs3"); print({k:v for k,v in t.items() if not k.startswith("stage")})
PY
ls gpurun_out | grep ${T}
