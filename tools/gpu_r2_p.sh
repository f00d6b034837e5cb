#!/bin/bash
# round 2, call P (1 GPU): mbarrier cluster barrier in the resident solve: tests that use it + the key-frame stream
O=gpurun_out/r2; mkdir -p $O
timeout 900 python -m pytest tests/test_session.py tests/test_adapter.py tests/test_gpu_parity.py -m gpu -q -x --deselect tests/test_gpu_parity.py::test_c5_full_size_properties > $O/p_tests.log 2>&1
echo "tests rc=$?" >> $O/p_tests.log; tail -4 $O/p_tests.log
timeout 400 python bench.py --workload stream --steps 1 --warmup 1 > $O/p_stream.json 2> $O/p_stream.err
python - $O/p_stream.json <<'PY'
import json, sys
for ln in open(sys.argv[1]):
    if ln.startswith("{"):
        d = json.loads(ln)
        cpu = d.get("cpu_baseline", {}).get("ms_per_keyframe")
        print(sys.argv[1].split("/")[-1], "ms/keyframe %.2f cpu %s per_keyframe %s" % (d["ms_per_keyframe"], cpu, {k: round(v, 3) for k, v in d["per_keyframe"].items()}))
PY
