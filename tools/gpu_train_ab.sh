#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --maxfail=5 --tb=short -p no:cacheprovider -k "train" 2>&1 | tail -6
for v in 0 1; do
  AB_ASYNC_WGRAD=$v timeout 300 python bench.py --steps 20 --no-network --no-cpu-baseline > gpurun_out/train_$v.json 2>gpurun_out/train_$v.err
  tail -c 300 gpurun_out/train_$v.err
  python - <<PY
import json
d=json.load(open("gpurun_out/train_$v.json"))
t=d["extras"]["train_loop_configs3"]
print("async_wgrad=$v", {k:(round(x,3) if isinstance(x,float) else x) for k,x in t.items() if not k.startswith("stage")})
PY
done
