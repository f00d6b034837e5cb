#!/bin/bash
# round 2, call X (1 GPU): the coarse path after the stale-plan fix: re-used handle (three passes of the stream), its tests,
# memcheck of a short three-pass stream, then the stream / C2-prefix numbers with and without it and the full GPU suite with it on
O=gpurun_out/r2; mkdir -p $O
CUDA_LAUNCH_BLOCKING=1 timeout 90 python tools/coarse_repro.py 400 3 > $O/x_native_repro.log 2>&1
echo "native repro rc=$?"; tail -4 $O/x_native_repro.log | cut -c1-300
timeout 150 python -m pytest tests/test_gpu_coarse.py -m gpu -q > $O/x_coarse_tests.log 2>&1
echo "coarse tests rc=$?"; tail -6 $O/x_coarse_tests.log | cut -c1-300
run() {  # name, workload, env...
  name=$1; wl=$2; shift; shift
  env "$@" timeout 120 python bench.py --workload $wl --steps 2 --warmup 1 --no-cpu-baseline > $O/x_$name.json 2> $O/x_$name.err
  echo "bench $name rc=$?"
  python - <<PY
import json
for ln in open("$O/x_$name.json"):
    if ln.startswith("{"):
        d = json.loads(ln)
        if "per_keyframe" in d: print("   ms/keyframe %.2f" % d["ms_per_keyframe"], {k: round(v, 3) for k, v in d["per_keyframe"].items()})
PY
}
run stream_coarse40 stream SGB_COARSE=1
run stream_coarse24 stream SGB_COARSE=1 SGB_COARSE_NODES=24
run stream_coarse16 stream SGB_COARSE=1 SGB_COARSE_NODES=16
timeout 120 compute-sanitizer --tool memcheck --print-limit 6 python tools/coarse_repro.py 60 3 > $O/x_memcheck_repro.log 2>&1
echo "memcheck repro rc=$?"; grep -E "^pass|Invalid|ERROR SUMMARY" $O/x_memcheck_repro.log | head -12 | cut -c1-260
SGB_COARSE=1 timeout 200 python -m pytest tests -m gpu -x -q > $O/x_all_tests_coarse_on.log 2>&1
echo "all gpu tests with SGB_COARSE=1 rc=$?" >> $O/x_all_tests_coarse_on.log; tail -3 $O/x_all_tests_coarse_on.log | cut -c1-300
