#!/bin/bash
# list-build sensitivity to the fat-group threshold (diagnostic)
for f in 0.9 1.2 1.5 2.0; do
  B2_FAT_FACTOR=$f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('c2 fat',$f, d['value'], d['engine']['list_stats'])"
done
for f in 0.9 1.5; do
  B2_FAT_FACTOR=$f python bench.py --reps 10 --steps 2 --warmup 3 --md-steps 20 --no-cpu-baseline --no-e2e | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('r10 fat',$f, d['value'], d['engine']['list_stats'])"
done
