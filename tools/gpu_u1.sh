#!/bin/bash
# quick: synthesis tests, bench line without the train legs, fresh per-kernel profile of one training step
set -u
mkdir -p gpurun_out
T=${1:-u1}
timeout 600 python -m pytest tests -m gpu -q --maxfail=12 --tb=short -p no:cacheprovider -k "synthesis or raster" > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/${T}_pytest.log
timeout 300 python bench.py --no-train --no-network --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
echo "bench rc=$?"; tail -c 400 gpurun_out/${T}_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/${T}_bench.json"))
print("views/s", d["value"], "e2e", d["e2e"]["value"], d["roofline"]["stage_ms_per_step"])
print(d["extras"]["synthesis_configs1"])
PY
timeout 300 python tools/prof_step_kernels.py > gpurun_out/${T}_step_kernels.txt 2>&1; head -70 gpurun_out/${T}_step_kernels.txt | cut -c1-150
