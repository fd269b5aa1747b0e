#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/n2_bench.json 2> gpurun_out/n2_bench.err
echo "rc=$?"; tail -c 500 gpurun_out/n2_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/n2_bench.json"))
ex=d.pop("extras")
print({k:d[k] for k in ("value","n_gpus","ms_per_step","scaling")}, "e2e", d["e2e"]["value"])
t=ex.get("train_loop_configs3"); print({k:v for k,v in t.items() if not k.startswith("stage")})
PY
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 | cut -c1-300
