#!/bin/bash
# round 2, call F (1 GPU): cluster-resident solve: GPU test tier, stream / C1 / C4 / C5 lines, sanitizer runs
O=gpurun_out/r2; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -rs --durations=8 > $O/f_tests.log 2>&1
echo "tests rc=$?" >> $O/f_tests.log; tail -25 $O/f_tests.log
timeout 400 python bench.py --workload stream --steps 1 --warmup 1 > $O/f_stream.json 2> $O/f_stream.err
python - <<'PY'
import json
for ln in open("gpurun_out/r2/f_stream.json"):
    if ln.startswith("{"):
        d = json.loads(ln)
        print("stream: ms/keyframe %.2f  cpu %.2f  per_keyframe %s" % (d["ms_per_keyframe"], d["cpu_baseline"]["ms_per_keyframe"], {k: round(v, 3) for k, v in d["per_keyframe"].items()}))
PY
SGB_NO_RESIDENT=1 timeout 400 python bench.py --workload stream --steps 1 --warmup 1 --no-cpu-baseline > $O/f_stream_nores.json 2> $O/f_stream_nores.err
python tools/show_line.py $O/f_stream_nores.json
for wl in c1 c2 c3 c4; do
  timeout 300 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline > $O/f_$wl.json 2> $O/f_$wl.err
  python tools/show_line.py $O/f_$wl.json
done
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/f_c5.json 2> $O/f_c5.err
python tools/show_line.py $O/f_c5.json
for tool in memcheck racecheck; do
  for wl in c1 c5s stream; do
    timeout 600 compute-sanitizer --tool $tool --log-file $O/f_${tool}_$wl.log python tools/sanitize_run.py $wl 2 > $O/f_${tool}_$wl.out 2>&1
    echo "$tool $wl rc=$? : $(tail -1 $O/f_${tool}_$wl.out) | $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $O/f_${tool}_$wl.log)"
  done
done
