#!/bin/bash
# round 2, call T (1 GPU): the final one-GPU record: whole GPU test tier, smoke, every bench line, launch list and ncu captures
O=gpurun_out/r2; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -rs > $O/t_tests.log 2>&1
echo "tests rc=$?" >> $O/t_tests.log; tail -5 $O/t_tests.log
timeout 300 python __graft_entry__.py smoke > $O/t_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/t_smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 > $O/t_c5.json 2> $O/t_c5.err
python tools/show_line.py $O/t_c5.json
for wl in c1 c2 c3 c4; do
  timeout 300 python bench.py --workload $wl --steps 3 --warmup 3 > $O/t_$wl.json 2> $O/t_$wl.err
  python tools/show_line.py $O/t_$wl.json
done
timeout 400 python bench.py --workload stream --steps 1 --warmup 1 > $O/t_stream.json 2> $O/t_stream.err
python - $O/t_stream.json <<'PY'
import json, sys
for ln in open(sys.argv[1]):
    if ln.startswith("{"):
        d = json.loads(ln)
        print("stream ms/keyframe %.2f cpu %s per_keyframe %s" % (d["ms_per_keyframe"], d.get("cpu_baseline", {}).get("ms_per_keyframe"), {k: round(v, 3) for k, v in d["per_keyframe"].items()}))
PY
SGB_MIN_WARMUP=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/t_launches_c5.csv \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $O/t_ncu_launches.log 2>&1
SGB_MIN_WARMUP=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pcg -s 28 -c 1 -o $O/t_pcg_c5 \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $O/t_ncu_pcg.log 2>&1
SGB_MIN_WARMUP=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_lin_lm|k_lin_pose|k_setup_chunk" -s 6 -c 3 -o $O/t_lin_c5 \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $O/t_ncu_lin.log 2>&1
ls -la $O/t_pcg_c5.ncu-rep $O/t_lin_c5.ncu-rep
timeout 400 python bench.py --impl reference --steps 1 --warmup 1 > $O/t_ref_c5.json 2> $O/t_ref_c5.err; tail -c 400 $O/t_ref_c5.json
