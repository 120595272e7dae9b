#!/bin/bash
# marginal cost of each level: FMG solve time of `L 8` for L = 7..3 (the coarse part of `7 8` is the `6 8` problem, ...)
for l in 7 6 5 4 3; do
  out=$(timeout 200 python bench.py --no-cpu-baseline --steps 20 --warmup 5 --log2-box-dim $l "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('%.3f ms/solve launches/solve %d' % (d['ms_per_step'], d['gpu_launches']/(d['steps']+d['warmup']) if d.get('gpu_launches') else -1))")
  echo "log2_box_dim=$l $@ :: $out"
done
