#!/bin/bash
# round 2, last call: the whole GPU suite on the final build, then per-launch times of the set-up / solve kernels on late key-frames
O=gpurun_out/r2; mkdir -p $O
timeout 125 python -m pytest tests -m gpu -q > $O/final_all_tests.log 2>&1
echo "all gpu tests, final build rc=$?" >> $O/final_all_tests.log; tail -3 $O/final_all_tests.log | cut -c1-300
timeout 40 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_setup_coarse|k_pcg_res4|k_setup_chunk" --launch-skip 14000 -c 450 --csv \
  --log-file $O/final_launches_stream_late.csv python tools/coarse_repro.py 330 1 > $O/final_ncu.log 2>&1
echo "ncu rc=$?"; python - <<'PY'
import csv, collections, io
try:
    rows = [l for l in open("gpurun_out/r2/final_launches_stream_late.csv") if l.startswith('"')]
    agg = collections.defaultdict(list)
    for r in csv.DictReader(io.StringIO("".join(rows))):
        try: agg[r["Kernel Name"][:40]].append(float(r["Metric Value"].replace(",", "")))
        except Exception: pass
    for k, v in agg.items(): print("   %-40s n=%d mean %.1f us max %.1f us" % (k, len(v), sum(v) / len(v) / 1e3, max(v) / 1e3))
except Exception as e:
    print("no launch list:", e)
PY
