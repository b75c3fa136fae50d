#!/usr/bin/env bash
# 8-GPU box: the driver's scaling run in miniature (N = 8 and 4; N = 1, 2 were measured on their own boxes).
mkdir -p gpurun_out
TAG="${1:-r01k}"
for N in 8 4; do
  MVD_BENCH_TRACE=150 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N \
    bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/${TAG}_bench_g$N.json 2> gpurun_out/${TAG}_bench_g$N.err
  echo "bench$N rc=$?"; cut -c1-420 gpurun_out/${TAG}_bench_g$N.json; grep -E "rank 0\]|Error|error" gpurun_out/${TAG}_bench_g$N.err | tail -4
done
