#!/bin/bash
set -u
mkdir -p gpurun_out
T=${1:-r1f}
timeout 900 python -m pytest tests -m gpu -q --maxfail=12 --tb=short -p no:cacheprovider -k "${2:-refine}" > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?"; tail -25 gpurun_out/${T}_pytest.log
timeout 600 python bench.py --no-train > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
echo "bench rc=$?"; tail -c 600 gpurun_out/${T}_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/${T}_bench.json"))
ex=d.pop("extras")
print(d["value"], d["e2e"]["value"])
print(json.dumps(ex.get("hand_obj_refiner_8f3")))
PY
