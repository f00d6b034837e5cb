#!/bin/bash
# usage: tools/bench_env.sh <workload> "<ENV=val ...>"...   -- one short bench line per environment setting (tuning runs)
wl=$1; shift
for v in "$@"; do
  echo "== $wl $v"
  env $v python bench.py --workload $wl --steps 2 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    ln = ln.strip()
    if not ln.startswith('{'):
        print(ln); continue
    d = json.loads(ln)
    r = d['roofline']
    print('value %.2f it/s  pcg %.0f GB/s frac %.3f  iters %d  phase_us %s  phases %s' % (d['value'], r['achieved'], r['frac'], r['pcg_iterations'], ['%.1f' % x for x in r['pcg_phase_us_per_iteration']], {k: round(v, 1) for k, v in d['phases_ms'].items()}))
"
done
