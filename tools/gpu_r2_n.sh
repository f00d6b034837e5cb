#!/bin/bash
# round 2, call N (1 GPU): ncu --set full of one k_pcg_res4 launch of the key-frame stream (late frame)
O=gpurun_out/r2; mkdir -p $O
SGB_MIN_WARMUP=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pcg_res4 -s 4500 -c 1 -o $O/n_res4_stream \
  python bench.py --workload stream --steps 1 --warmup 0 --no-cpu-baseline --stream-frames 320 > $O/n_ncu_res4.log 2>&1
ls -la $O/n_res4_stream.ncu-rep; tail -3 $O/n_ncu_res4.log
