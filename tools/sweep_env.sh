#!/bin/bash
# usage: tools/sweep_env.sh "VAR=val VAR2=val" ...   -> one bench line (ms/solve, us/sweep, norm) per setting
for cfg in "$@"; do
  out=$(env $cfg timeout 200 python bench.py --no-cpu-baseline --steps 10 --warmup 3 $BENCH_ARGS 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('%.3f ms/solve  %.1f us/sweep  norm %.15e' % (d['ms_per_step'], d['roofline']['us_per_launch'], d['f_cycle_norm']))")
  echo "$cfg :: $out"
done
