#!/bin/bash
# round 2, call B (2 GPUs): multi-GPU parity tests in the three planner layouts, C5 at N=2 (default = ghost rows +
# rank-filtered symbolic phase, then the owner-only layout), N=1 on the same box for the efficiency ratio
O=gpurun_out/r2; mkdir -p $O
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -q -rs > $O/b_multigpu_tests.log 2>&1
echo "tests rc=$?" >> $O/b_multigpu_tests.log; tail -4 $O/b_multigpu_tests.log
run2() {  # name, env...
  name=$1; shift
  env "$@" SGB_PROFILE=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
    --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > $O/b_c5_n2_$name.json 2> $O/b_c5_n2_$name.err
  echo "bench $name rc=$?"; tail -c 900 $O/b_c5_n2_$name.json
}
run2 default
run2 owner_only SGB_GHOST_LANDMARKS=0
run2 ghost_full SGB_PARTITION_FULL=1
timeout 400 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline > $O/b_c5_n1.json 2> $O/b_c5_n1.err
echo "bench n1 rc=$?"; tail -c 600 $O/b_c5_n1.json
