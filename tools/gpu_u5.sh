#!/bin/bash
set -u
mkdir -p gpurun_out
T=${1:-u5}
timeout 1200 python -m pytest tests -m gpu -q --maxfail=12 --tb=short -p no:cacheprovider > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/${T}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
echo "bench rc=$?"; tail -c 400 gpurun_out/${T}_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/${T}_bench.json"))
ex=d.pop("extras")
print(json.dumps({k:d[k] for k in ("value","ms_per_step","e2e","gpu_launches","clocks","wall_s")}))
print(json.dumps({k:v for k,v in d["roofline"].items() if k!="note"}))
print({k:v for k,v in d.items() if k.startswith("train") or k.startswith("synth") or k.startswith("mpcpe")})
t=ex.get("train_loop_configs3"); print({k:v for k,v in t.items() if not k.startswith("stage")})
print(ex.get("synthesis_configs1"))
PY
