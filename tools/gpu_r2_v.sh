#!/bin/bash
# round 2, call V (8 GPUs): final multi-GPU record with the final build: parity tests at 8 and 4 ranks, C5 at N=8 (default,
# then with the former 2-step landmark slices) and N=4
O=gpurun_out/r2; mkdir -p $O
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -q -rs -k "8-ghost-rows-filtered or 4-ghost-rows-filtered or 8-owner-only" > $O/v_multigpu_tests.log 2>&1
echo "tests rc=$?" >> $O/v_multigpu_tests.log; tail -3 $O/v_multigpu_tests.log
runN() {  # n, name, env...
  n=$1; name=$2; shift; shift
  env "$@" SGB_PROFILE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 \
    --master-port 29519 bench.py --gpus $n --steps 3 --warmup 3 --no-cpu-baseline > $O/v_c5_n${n}_$name.json 2> $O/v_c5_n${n}_$name.err
  echo "bench n=$n $name rc=$?"; python tools/show_line.py $O/v_c5_n${n}_$name.json; grep -m1 "pcg grid" $O/v_c5_n${n}_$name.err
}
runN 8 default
runN 8 lmsteps2 SGB_LM_STEPS=2
runN 4 default
