#!/bin/bash
# round 2, call A (2 GPUs): the multi-GPU parity tests on hardware (both landmark modes), C5 at N=2 with and without
# ghost landmarks (bench carries parity_vs_n1), memcheck of one 2-rank run
O=gpurun_out/r2; mkdir -p $O
nvidia-smi -L > $O/a_gpus.txt 2>&1
nvidia-smi topo -m >> $O/a_gpus.txt 2>&1
timeout 1200 python -m pytest tests/test_multigpu.py -m gpu -q -rs > $O/a_multigpu_tests.log 2>&1
echo "tests rc=$?" >> $O/a_multigpu_tests.log
for gh in 0 1; do
  SGB_GHOST_LANDMARKS=$gh timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
    --master-port 2951$gh bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > $O/a_c5_n2_ghost$gh.json 2> $O/a_c5_n2_ghost$gh.err
  echo "bench ghost=$gh rc=$?"; tail -c 600 $O/a_c5_n2_ghost$gh.json
done
timeout 600 compute-sanitizer --tool memcheck --target-processes all --log-file $O/a_memcheck_n2.%p.log \
  python -m pytest tests/test_multigpu.py -m gpu -q -k "2-0" > $O/a_memcheck_n2_pytest.log 2>&1
echo "memcheck rc=$?"
tail -5 $O/a_multigpu_tests.log
