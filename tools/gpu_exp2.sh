#!/bin/bash
# A/B of raster variants on the bench's ten resident batches: usage gpu_exp2.sh "name1 name2 ..." ("orig" = the shipped library)
cp artiboost_b200/libartiboost_b200.so /tmp/lib_orig.so
for v in $1; do
  if [ "$v" = orig ]; then cp /tmp/lib_orig.so artiboost_b200/libartiboost_b200.so; else cp artiboost_b200/build/variants/$v.so artiboost_b200/libartiboost_b200.so; fi
  timeout 300 python bench.py --steps 100 --warmup 5 --no-train --no-network --no-cpu-baseline 2>/dev/null | tail -1 > /tmp/line.json
  python -c "
import json
d=json.load(open('/tmp/line.json'))
print('$v', round(d['value']), round(d['ms_per_step'],4), d['roofline']['stage_ms_per_step'])"
done
cp /tmp/lib_orig.so artiboost_b200/libartiboost_b200.so
