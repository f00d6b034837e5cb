#!/bin/bash
# round 2, call M (4 GPUs): block-size variants of k_pcg at N=4 (2.2 rows per thread with 256 threads)
O=gpurun_out/r2; mkdir -p $O
runN() {  # n, name, env...
  n=$1; name=$2; shift; shift
  env "$@" SGB_PROFILE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 \
    --master-port 29519 bench.py --gpus $n --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $O/m_c5_n${n}_$name.json 2> $O/m_c5_n${n}_$name.err
  echo "bench n=$n $name rc=$?"; python tools/show_line.py $O/m_c5_n${n}_$name.json; grep -m1 "pcg grid" $O/m_c5_n${n}_$name.err
}
runN 4 bt288 SGB_PCG_THREADS=288
runN 4 bt320 SGB_PCG_THREADS=320
