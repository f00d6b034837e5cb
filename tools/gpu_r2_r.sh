#!/bin/bash
# round 2, call R (1 GPU): lane-per-block landmark linearisation: the whole GPU test tier, launch list and bench line of C5
O=gpurun_out/r2; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x -rs > $O/r_tests.log 2>&1
echo "tests rc=$?" >> $O/r_tests.log; tail -6 $O/r_tests.log
SGB_MIN_WARMUP=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r_launches_c5.csv \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $O/r_ncu_launches.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r2/r_launches_c5.csv")) if len(r) > 10 and r[0].isdigit()]
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    name = r[4].split("(")[0][:60]
    agg[name][0] += 1
    agg[name][1] += float(r[-1].replace(",", "")) / 1e3
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:9]:
    print("%-40s n=%4d total %10.1f us  mean %9.1f us  %5.1f %%" % (k, v[0], v[1], v[1] / v[0], 100 * v[1] / tot))
PY
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/r_c5.json 2> $O/r_c5.err
python tools/show_line.py $O/r_c5.json
