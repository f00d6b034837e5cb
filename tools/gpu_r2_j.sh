#!/bin/bash
# round 2, call J (2 GPUs): N=2 with the block-size variants of k_pcg, recycled arenas, threaded global scans
O=gpurun_out/r2; mkdir -p $O
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -q -rs -k "2-" > $O/j_multigpu_tests.log 2>&1
echo "tests rc=$?" >> $O/j_multigpu_tests.log; tail -3 $O/j_multigpu_tests.log
run2() {  # name, env...
  name=$1; shift
  env "$@" SGB_PROFILE=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
    --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > $O/j_c5_n2_$name.json 2> $O/j_c5_n2_$name.err
  echo "bench $name rc=$?"; python tools/show_line.py $O/j_c5_n2_$name.json; grep -m1 "pcg grid" $O/j_c5_n2_$name.err
}
run2 default
run2 bt288 SGB_PCG_THREADS=288
run2 bt320 SGB_PCG_THREADS=320
