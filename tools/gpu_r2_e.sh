#!/bin/bash
# round 2, call E (1 GPU): the whole GPU test tier, block-size variants of k_pcg on C5, launch list of a C5 step,
# the key-frame stream with its per-call breakdown, the small configs
O=gpurun_out/r2; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x -rs --durations=15 > $O/e_tests.log 2>&1
echo "tests rc=$?" >> $O/e_tests.log; tail -30 $O/e_tests.log
for bt in 256 288 320; do
  SGB_PCG_THREADS=$bt timeout 300 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $O/e_c5_bt$bt.json 2> $O/e_c5_bt$bt.err
  python tools/show_line.py $O/e_c5_bt$bt.json
done
SGB_MIN_WARMUP=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/e_launches_c5.csv \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $O/e_ncu_launches.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r2/e_launches_c5.csv")) if len(r) > 10 and r[0].isdigit()]
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    name = r[4].split("(")[0][:60]
    agg[name][0] += 1
    agg[name][1] += float(r[-1].replace(",", "")) / 1e3
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-62s n=%4d total %10.1f us  mean %9.1f us  %5.1f %%" % (k, v[0], v[1], v[1] / v[0], 100 * v[1] / tot))
PY
timeout 400 python bench.py --workload stream --steps 1 --warmup 1 > $O/e_stream.json 2> $O/e_stream.err
python - <<'PY'
import json
for ln in open("gpurun_out/r2/e_stream.json"):
    if ln.startswith("{"):
        d = json.loads(ln)
        print("stream: ms/keyframe %.2f  cpu %.2f  per_keyframe %s" % (d["ms_per_keyframe"], d["cpu_baseline"]["ms_keyframe"] if "ms_keyframe" in d["cpu_baseline"] else d["cpu_baseline"]["ms_per_keyframe"], {k: round(v, 3) for k, v in d["per_keyframe"].items()}))
PY
for wl in c1 c2 c3; do
  timeout 300 python bench.py --workload $wl --steps 3 --warmup 3 > $O/e_$wl.json 2> $O/e_$wl.err
  python tools/show_line.py $O/e_$wl.json
done
