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
run c1_coarse c1 SGB_COARSE=1
run c1_plain c1 SGB_COARSE=0
for tool in racecheck memcheck; do
  timeout 150 compute-sanitizer --tool $tool python tools/sanitize_run.py coarse 2 > $O/w_${tool}_coarse.log 2>&1
  echo "$tool coarse rc=$? : $(grep -c 'coarse .* ok' $O/w_${tool}_coarse.log) runs ok | $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $O/w_${tool}_coarse.log | tail -1)"
done
SGB_COARSE=1 timeout 240 python -m pytest tests -m gpu -x -q > $O/w_all_tests_coarse_on.log 2>&1
echo "all gpu tests with SGB_COARSE=1 rc=$?" >> $O/w_all_tests_coarse_on.log; tail -4 $O/w_all_tests_coarse_on.log
