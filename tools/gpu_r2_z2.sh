#!/bin/bash
# round 2, call Z2 (the last seconds of the budget): the two-level preconditioner inside the batched kernel (k_lm_block)
O=gpurun_out/r2; mkdir -p $O
timeout 60 python -m pytest tests/test_gpu_parity.py -m gpu -q -k batched > $O/z2_batch_tests_on.log 2>&1; echo "batch tests (coarse on) rc=$?"; tail -2 $O/z2_batch_tests_on.log | cut -c1-200
timeout 50 python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline > $O/z2_c4_on.json 2> $O/z2_c4_on.err; echo "c4 on rc=$?"; python tools/show_line.py $O/z2_c4_on.json | cut -c1-120
SGB_COARSE_BATCH=0 timeout 50 python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline > $O/z2_c4_off.json 2> $O/z2_c4_off.err; echo "c4 off rc=$?"; python tools/show_line.py $O/z2_c4_off.json | cut -c1-120
SGB_COARSE_BATCH=0 timeout 60 python -m pytest tests/test_gpu_parity.py -m gpu -q -k batched > $O/z2_batch_tests_off.log 2>&1; echo "batch tests (coarse off) rc=$?"; tail -2 $O/z2_batch_tests_off.log | cut -c1-200
python - <<'PY'
import json
for n in ("on", "off"):
    for ln in open("gpurun_out/r2/z2_c4_%s.json" % n):
        if ln.startswith("{"):
            d = json.loads(ln); print(n, "value", round(d["value"]), "ms/step", round(d["ms_per_step"], 2), {k: d[k] for k in d if "pcg" in k or "iters" in k})
PY
