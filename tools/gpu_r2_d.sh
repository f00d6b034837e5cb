#!/bin/bash
# round 2, call D (2 GPUs): pushed pose halos + recycled arenas: parity tests, C5 at N=2 (default / pull halos), N=1
O=gpurun_out/r2; mkdir -p $O
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -q -rs -k "2-" > $O/d_multigpu_tests.log 2>&1
echo "tests rc=$?" >> $O/d_multigpu_tests.log; tail -4 $O/d_multigpu_tests.log
run2() {  # name, env...
  name=$1; shift
  env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
    --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > $O/d_c5_n2_$name.json 2> $O/d_c5_n2_$name.err
  echo "bench $name rc=$?"; python tools/show_line.py $O/d_c5_n2_$name.json
}
run2 default
run2 pull_halos SGB_PUSHED_HALOS=0
timeout 400 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline > $O/d_c5_n1.json 2> $O/d_c5_n1.err
echo "bench n1 rc=$?"; python tools/show_line.py $O/d_c5_n1.json
