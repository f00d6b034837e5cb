#!/bin/bash
# round 2, call G (1 GPU): single-CTA / cluster resident solves, batched kernel with the resident solve, online path
O=gpurun_out/r2; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -rs -x --deselect tests/test_gpu_parity.py::test_c5_full_size_properties > $O/g_tests.log 2>&1
echo "tests rc=$?" >> $O/g_tests.log; tail -12 $O/g_tests.log
stream() {  # name, env...
  name=$1; shift
  env "$@" timeout 400 python bench.py --workload stream --steps 1 --warmup 1 $CPUFLAG > $O/g_stream_$name.json 2> $O/g_stream_$name.err
  python - $O/g_stream_$name.json <<'PY'
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
stream cluster_only SGB_NO_RESIDENT1=1
stream no_online SGB_NO_RESIDENT1=0 SGB_SESSION_FULL=1
SGB_PROFILE=1 timeout 300 python bench.py --workload c1 --steps 3 --warmup 3 --no-cpu-baseline > $O/g_c1.json 2> $O/g_c1.err
python tools/show_line.py $O/g_c1.json; grep -m2 "resident" $O/g_c1.err
timeout 300 python bench.py --workload c4 --steps 3 --warmup 3 > $O/g_c4.json 2> $O/g_c4.err
python tools/show_line.py $O/g_c4.json
SGB_MIN_WARMUP=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pcg_res -s 3000 -c 2 -o $O/g_res_stream \
  python bench.py --workload stream --steps 1 --warmup 0 --no-cpu-baseline --stream-frames 320 > $O/g_ncu_res.log 2>&1
ls -la $O/g_res_stream.ncu-rep
