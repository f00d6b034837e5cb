#!/bin/bash
# round 2, call Y (1 GPU, the last one): the whole GPU suite with the two-level preconditioner ON (the new default) and OFF,
# the key-frame stream bench line of record (with its CPU arm), smoke()
O=gpurun_out/r2; mkdir -p $O
timeout 170 python -m pytest tests -m gpu -q > $O/y_all_tests_default.log 2>&1
echo "all gpu tests, default (coarse on) rc=$?" >> $O/y_all_tests_default.log; tail -3 $O/y_all_tests_default.log | cut -c1-300
timeout 100 python bench.py --workload stream --steps 2 --warmup 1 > $O/y_stream.json 2> $O/y_stream.err
echo "bench stream rc=$?"
python - <<'PY'
import json
for ln in open("gpurun_out/r2/y_stream.json"):
    if ln.startswith("{"):
        d = json.loads(ln)
        print("   ms/keyframe %.2f cpu %.2f" % (d["ms_per_keyframe"], d.get("cpu_baseline", {}).get("ms_per_keyframe", 0)), {k: round(v, 3) for k, v in d["per_keyframe"].items()})
PY
SGB_COARSE=0 timeout 170 python -m pytest tests -m gpu -q > $O/y_all_tests_coarse_off.log 2>&1
echo "all gpu tests, SGB_COARSE=0 rc=$?" >> $O/y_all_tests_coarse_off.log; tail -3 $O/y_all_tests_coarse_off.log | cut -c1-300
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
