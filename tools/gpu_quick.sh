#!/bin/bash
# Quick GPU iteration: selected parity tests + the raster-only bench line (stage split).  usage: gpu_quick.sh TAG [pytest -k expr]
set -u
mkdir -p gpurun_out
T=${1:-q}
K=${2:-"raster or refine or synthesis"}
timeout 900 python -m pytest tests -m gpu -q --maxfail=12 --tb=short -p no:cacheprovider -k "$K" > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?"; tail -25 gpurun_out/${T}_pytest.log
timeout 300 python bench.py --no-train --no-network --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
echo "bench rc=$?"; tail -c 400 gpurun_out/${T}_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/${T}_bench.json"))
print("views/s", d["value"], "e2e", d["e2e"]["value"], d["roofline"]["stage_ms_per_step"], "frac", d["roofline"]["frac"], d["roofline"]["whole_step"])
PY
