#!/bin/bash
set -u
mkdir -p gpurun_out
T=${1:-r1d}
timeout 900 python -m pytest tests -m gpu -q --maxfail=12 --tb=short -p no:cacheprovider > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/${T}_pytest.log
for mb in 4 5 6; do
  AB_TRI_MINB=$mb timeout 200 python bench.py --steps 100 --no-train --no-network --no-cpu-baseline > gpurun_out/var.json 2>/dev/null
  python - <<PY
import json
d=json.load(open("gpurun_out/var.json"))
print("minb=$mb views/s %.0f" % d["value"], {k: round(x,4) for k,x in d["roofline"]["stage_ms_per_step"].items()})
PY
done
timeout 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
echo "bench rc=$?"; tail -c 600 gpurun_out/${T}_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/${T}_bench.json"))
ex=d.pop("extras")
print(json.dumps({k:d[k] for k in ("value","ms_per_step","e2e","gpu_launches","roofline","cpu_baseline","clocks")}))
print(json.dumps(ex.get("hand_obj_refiner_8f3")))
t=ex.get("train_loop_configs3"); print({k:v for k,v in t.items() if not k.startswith("stage")})
PY
