#!/bin/bash
# round 2, call I (1 GPU): four-lane resident solve: tests, stream, C1, C4
O=gpurun_out/r2; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -rs -x --deselect tests/test_gpu_parity.py::test_c5_full_size_properties > $O/i_tests.log 2>&1
echo "tests rc=$?" >> $O/i_tests.log; tail -12 $O/i_tests.log
stream() {  # name, env...
  name=$1; shift
  env "$@" timeout 400 python bench.py --workload stream --steps 1 --warmup 1 $CPUFLAG > $O/i_stream_$name.json 2> $O/i_stream_$name.err
  python - $O/i_stream_$name.json <<'PY'
import json, sys
for ln in open(sys.argv[1]):
    if ln.startswith("{"):
        d = json.loads(ln)
        cpu = d.get("cpu_baseline", {}).get("ms_per_keyframe")
        print(sys.argv[1].split("/")[-1], "ms/keyframe %.2f cpu %s per_keyframe %s" % (d["ms_per_keyframe"], cpu, {k: round(v, 3) for k, v in d["per_keyframe"].items()}))
PY
}
CPUFLAG=""; stream default
CPUFLAG="--no-cpu-baseline"
stream no_res4 SGB_NO_RES4=1
for wl in c1 c3; do
SGB_PROFILE=1 timeout 300 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline > $O/i_$wl.json 2> $O/i_$wl.err
python tools/show_line.py $O/i_$wl.json; grep -m2 "four lanes" $O/i_$wl.err
done
timeout 300 python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline > $O/i_c4.json 2> $O/i_c4.err
python tools/show_line.py $O/i_c4.json
