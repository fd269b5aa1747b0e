#!/bin/bash
# Final verification + evidence of round 2: all GPU tests, smoke(), the bench line + reference arm, the ncu launch list of the
# same bench command, full captures of the raster kernels and of the training-step kernels this round rewrote.
set -u
mkdir -p gpurun_out
T=${1:-r2f}
timeout 1200 python -m pytest tests -m gpu -q --maxfail=12 --tb=short -p no:cacheprovider > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/${T}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
echo "bench rc=$?"; tail -c 300 gpurun_out/${T}_bench.err
timeout 400 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${T}_bench_ref.json 2>> gpurun_out/${T}_bench.err
echo "ref rc=$?"
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-network --no-train"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_bench.csv $B > gpurun_out/${T}_bench_under_ncu.log 2>&1
echo "launch list rc=$? lines=$(wc -l < gpurun_out/${T}_launches_bench.csv)"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"raster_(bin|tile)_kernel" -s 8 -c 2 -f -o gpurun_out/${T}_raster $B > gpurun_out/${T}_ncu_raster.log 2>&1
echo "raster capture rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gemm_bf16_tn_persistent_kernel" -s 6 -c 3 -f -o gpurun_out/${T}_gemm_persistent python tools/time_gemm.py > gpurun_out/${T}_ncu_gemm.log 2>&1
echo "gemm capture rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"wgrad_bf16_kernel" -s 4 -c 4 -f -o gpurun_out/${T}_wgrad python tools/time_wgrad.py > gpurun_out/${T}_ncu_wgrad.log 2>&1
echo "wgrad capture rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"bn_bwd_(reduce|apply)_kernel|im2col_c4_pairs" -s 20 -c 4 -f -o gpurun_out/${T}_bn python tools/time_bn.py > gpurun_out/${T}_ncu_bn.log 2>&1
echo "bn capture rc=$?"
timeout 300 python tools/prof_step_kernels.py > gpurun_out/${T}_train_step_kernels.txt 2>&1
echo "step kernels rc=$?"; sed -n 3,4p gpurun_out/${T}_train_step_kernels.txt
python tools/time_conv.py > gpurun_out/${T}_conv.txt 2>&1; python tools/time_gemm.py > gpurun_out/${T}_gemm.txt 2>&1; python tools/time_wgrad.py > gpurun_out/${T}_wgrad.txt 2>&1; python tools/time_bn.py > gpurun_out/${T}_bn.txt 2>&1
python tools/time_train_step.py | tail -1 | tee gpurun_out/${T}_step.txt
python - <<PY
import json
d=json.load(open("gpurun_out/${T}_bench.json"))
ex=d.pop("extras")
print(json.dumps({k:d[k] for k in ("value","ms_per_step","e2e","gpu_launches","clocks","wall_s","oracle_checked")}))
print(json.dumps({k:v for k,v in d["roofline"].items() if k!="note"}))
print({k:v for k,v in d.items() if k.startswith("train") or k.startswith("synth") or k.startswith("mpcpe")})
t=ex.get("train_loop_configs3"); print({k:v for k,v in t.items() if not k.startswith("stage")})
print({k:{kk:vv for kk,vv in v.items() if not kk.startswith("stage")} for k,v in ex.get("network_forward_configs2",{}).items()})
print(open("gpurun_out/${T}_bench_ref.json").read()[:400])
PY
ls gpurun_out | grep ${T}
