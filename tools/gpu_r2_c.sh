#!/bin/bash
# round 2, call C (8 GPUs): multi-GPU parity tests at 8 and 4 ranks, C5 at N=8 and N=4 (bench carries parity_vs_n1)
O=gpurun_out/r2; mkdir -p $O
nvidia-smi -L > $O/c_gpus.txt 2>&1
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -q -rs -k "8-ghost-rows-filtered or 8-owner-only or 4-ghost-rows-filtered or 8-ghost-rows" > $O/c_multigpu_tests.log 2>&1
echo "tests rc=$?" >> $O/c_multigpu_tests.log; tail -4 $O/c_multigpu_tests.log
runN() {  # n, name, env...
  n=$1; name=$2; shift; shift
  env "$@" SGB_PROFILE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 \
    --master-port 29519 bench.py --gpus $n --steps 3 --warmup 3 --no-cpu-baseline > $O/c_c5_n${n}_$name.json 2> $O/c_c5_n${n}_$name.err
  echo "bench n=$n $name rc=$?"; tail -c 1200 $O/c_c5_n${n}_$name.json
}
runN 8 default
runN 4 default
runN 8 owner_only SGB_GHOST_LANDMARKS=0
