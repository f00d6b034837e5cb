#!/bin/bash
# round 2, call W (1 GPU, last one): the two-level preconditioner of the resident solve (sgb_coarse.h) on hardware:
# its parity tests, the key-frame stream and C1 with / without it, racecheck + memcheck, then the whole GPU suite with it on.
O=gpurun_out/r2; mkdir -p $O
timeout 200 python -m pytest tests/test_gpu_coarse.py -m gpu -x -q > $O/w_coarse_tests.log 2>&1
echo "coarse tests rc=$?" >> $O/w_coarse_tests.log; tail -4 $O/w_coarse_tests.log
run() {  # name, workload, env...
  name=$1; wl=$2; shift; shift
  env "$@" timeout 120 python bench.py --workload $wl --steps 2 --warmup 1 --no-cpu-baseline > $O/w_$name.json 2> $O/w_$name.err
  echo "bench $name rc=$?"; python tools/show_line.py $O/w_$name.json
  python - <<PY
import json
for ln in open("$O/w_$name.json"):
    if ln.startswith("{"):
        d = json.loads(ln)
        if "per_keyframe" in d: print("   ms/keyframe %.2f" % d["ms_per_keyframe"], {k: round(v, 3) for k, v in d["per_keyframe"].items()})
PY
}
run stream_coarse40 stream SGB_COARSE=1
run stream_coarse24 stream SGB_COARSE=1 SGB_COARSE_NODES=24
run stream_plain stream SGB_COARSE=0
run c1_coarse c1 SGB_COARSE=1 SGB_PROFILE=1
grep -m1 'four lanes' $O/w_c1_coarse.err
run c1_plain c1 SGB_COARSE=0
for tool in racecheck memcheck; do
  timeout 150 compute-sanitizer --tool $tool python tools/sanitize_run.py coarse 2 > $O/w_${tool}_coarse.log 2>&1
  echo "$tool coarse rc=$? : $(grep -c 'coarse .* ok' $O/w_${tool}_coarse.log) runs ok | $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $O/w_${tool}_coarse.log | tail -1)"
done
SGB_COARSE=1 timeout 240 python -m pytest tests -m gpu -x -q > $O/w_all_tests_coarse_on.log 2>&1
echo "all gpu tests with SGB_COARSE=1 rc=$?" >> $O/w_all_tests_coarse_on.log; tail -4 $O/w_all_tests_coarse_on.log
SGB_COARSE=1 timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_setup_coarse|k_pcg_res4|k_setup_chunk" -c 900 --csv \
  --log-file $O/w_launches_stream_coarse.csv python tools/sanitize_run.py stream 15 > $O/w_ncu_stream.log 2>&1
echo "ncu rc=$?"; python - <<'PY'
import csv, collections, io
rows = [l for l in open("gpurun_out/r2/w_launches_stream_coarse.csv") if l.startswith('"')]
agg = collections.defaultdict(list)
for r in csv.DictReader(io.StringIO("".join(rows))):
    try: agg[r["Kernel Name"][:40]].append(float(r["Metric Value"].replace(",", "")))
    except Exception: pass
for k, v in agg.items(): print("   %-40s n=%d mean %.1f us max %.1f us" % (k, len(v), sum(v) / len(v) / 1e3, max(v) / 1e3))
PY
