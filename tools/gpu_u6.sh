#!/bin/bash
set -u
mkdir -p gpurun_out
T=u6
timeout 600 python -m pytest tests -m gpu -q --maxfail=12 --tb=short -p no:cacheprovider -k "column_statistics or trainstep or train_ops or wgrad or conv or mano or pose" 2>&1 | tail -4 | tee gpurun_out/${T}_pytest.log
python tools/time_train_step.py 2>&1 | tail -1 | tee gpurun_out/${T}_step.txt
for p in -1 0; do
  echo "== AB_LOOP_PRIO=$p" | tee -a gpurun_out/${T}_loop.txt
  AB_LOOP_PRIO=$p timeout 300 python bench.py --no-network --no-cpu-baseline --equiv-steps 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
t=d['extras']['train_loop_configs3']
print({k:t[k] for k in ('images_per_s','ms_per_step','ms_train_step_only','ms_synthesis_and_batching')})" | tee -a gpurun_out/${T}_loop.txt
done
cp artiboost_b200/libartiboost_b200.so /tmp/lib_orig.so
for v in base lbs44 lbs44b; do
  if [ $v != base ]; then cp artiboost_b200/build/variants/$v.so artiboost_b200/libartiboost_b200.so; fi
  echo "== $v" | tee -a gpurun_out/${T}_lbs.txt
  python tools/prof_lbs.py 2>&1 | grep mano | tee -a gpurun_out/${T}_lbs.txt
  if [ $v != base ]; then timeout 300 python -m pytest tests -m gpu -q -p no:cacheprovider -k "mano or pose_generator" 2>&1 | tail -1 | tee -a gpurun_out/${T}_lbs.txt; fi
done
cp /tmp/lib_orig.so artiboost_b200/libartiboost_b200.so
