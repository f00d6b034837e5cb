#!/bin/bash
# round 2, call S (2 GPUs): the lane-per-block landmark linearisation on the partitioned path (bit-identical Hessian across
# rank counts is asserted by the test) + the N=2 bench line of the final build
O=gpurun_out/r2; mkdir -p $O
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -q -rs -k "2-" > $O/s_multigpu_tests.log 2>&1
echo "tests rc=$?" >> $O/s_multigpu_tests.log; tail -3 $O/s_multigpu_tests.log
SGB_PROFILE=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
  --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > $O/s_c5_n2.json 2> $O/s_c5_n2.err
echo "bench rc=$?"; python tools/show_line.py $O/s_c5_n2.json
