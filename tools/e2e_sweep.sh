#!/bin/bash
# e2e (pool) throughput for several slot/batch shapes; prints sparse and dense fps
for cfg in "16 2 64" "24 2 80" "16 3 80" "32 1 80"; do
  set -- $cfg
  timeout 120 python bench.py --no-cpu-baseline --steps 5 --warmup 3 --pool-slots $1 --pool-batch $2 --pool-inflight $3 2>/dev/null | \
    python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$cfg', round(d['e2e']['value']), round(d['e2e']['dense_format']['value']))"
done
