#!/bin/bash
for o in 0 3 7 20; do
  AB_BENCH_SEED_OFFSET=$o timeout 300 python bench.py --steps 20 --warmup 5 --no-train --no-network --no-cpu-baseline 2>/dev/null | tail -1 > /tmp/line.json
  python -c "
import json
d=json.load(open('/tmp/line.json'))
print('offset $o', round(d['value']), [round(x,4) for x in d['per_rank_ms']], d['step_ms_distribution'])"
done
