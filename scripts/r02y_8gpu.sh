#!/usr/bin/env bash
# 8-GPU box, final state of round 2: parity at 8 ranks, scaling points N = 8 and 4 (multicast-epilogue gathers), MultiviewX on
# 6 of 8 GPUs, per-kernel timeline of one 8-GPU step.
mkdir -p gpurun_out
TAG="${1:-r02y}"
echo "== sharded check 8 ranks"; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 scripts/check_sharded.py > gpurun_out/${TAG}_sharded8.log 2>&1; echo "check rc=$?"; grep -E "SHARDED|Error|error" gpurun_out/${TAG}_sharded8.log | tail -4
for CFG in "8 wildtrack" "4 wildtrack" "8 multiviewx" "8 stress4k"; do set -- $CFG
  MVD_BENCH_TRACE=200 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 2953$1 bench.py --gpus $1 --steps 20 --warmup 5 --workload $2 > gpurun_out/${TAG}_bench_g$1_$2.json 2> gpurun_out/${TAG}_bench_g$1_$2.err
  echo "bench N=$1 $2 rc=$?"; cut -c1-170 gpurun_out/${TAG}_bench_g$1_$2.json; grep -E "rank 0\] timed|Error|error" gpurun_out/${TAG}_bench_g$1_$2.err | tail -2
done
echo "== timeline 8 GPUs"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 scripts/timeline.py --out gpurun_out/${TAG}_timeline_8gpu > /dev/null 2> gpurun_out/${TAG}_timeline_8gpu.err; echo "rc=$?"; head -14 gpurun_out/${TAG}_timeline_8gpu.txt | cut -c1-150
