#!/bin/bash
# round 2, call K (1 GPU): launch list of the key-frame stream (which kernels the 15 ms per key-frame are made of)
O=gpurun_out/r2; mkdir -p $O
SGB_MIN_WARMUP=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 30000 -c 700 --csv --log-file $O/k_launches_stream.csv \
  python bench.py --workload stream --steps 1 --warmup 0 --no-cpu-baseline --stream-frames 320 > $O/k_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r2/k_launches_stream.csv")) if len(r) > 10 and r[0].isdigit()]
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    name = r[4].split("(")[0][:60]
    agg[name][0] += 1
    agg[name][1] += float(r[-1].replace(",", "")) / 1e3
tot = sum(v[1] for v in agg.values())
print("launches", len(rows), "total us", tot)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-40s n=%4d total %10.1f us  mean %9.1f us  %5.1f %%" % (k, v[0], v[1], v[1] / v[0], 100 * v[1] / tot))
PY
