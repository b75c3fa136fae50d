#!/usr/bin/env bash
# 8-GPU box: parity of the view-sharded path at 8 ranks (one rank without views), the scaling points N = 8 and 4 with the
# fused multicast gathers and with ncclAllGather, BASELINE configs[2] (MultiviewX, 6 views on 6 of 8 GPUs) and configs[3]
# (8-view 4K stress), and the per-kernel timeline of one 8-GPU step.
mkdir -p gpurun_out
TAG="${1:-r02l}"
run() { N=$1; PORT=$2; shift 2; python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT "$@"; }
echo "== sharded check 8 ranks"; timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 scripts/check_sharded.py > gpurun_out/${TAG}_sharded8.log 2>&1; echo "check rc=$?"; grep -E "SHARDED|Error|error" gpurun_out/${TAG}_sharded8.log | tail -4; grep -E "^wildtrack|^multiviewx|^one_view" gpurun_out/${TAG}_sharded8.log | cut -c1-260 | tail -6
for CFG in "8 fused 1" "8 nccl 0" "4 fused 1"; do set -- $CFG
  MVDETR_B200_FUSED_GATHER=$3 MVD_BENCH_TRACE=150 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 2953$1 bench.py --gpus $1 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_g$1_$2.json 2> gpurun_out/${TAG}_bench_g$1_$2.err
  echo "bench N=$1 $2 rc=$?"; cut -c1-160 gpurun_out/${TAG}_bench_g$1_$2.json; grep -E "rank 0\]|Error|error" gpurun_out/${TAG}_bench_g$1_$2.err | tail -3
done
for WL in multiviewx stress4k; do
  MVD_BENCH_TRACE=200 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 10 --warmup 3 --workload $WL > gpurun_out/${TAG}_bench_g8_$WL.json 2> gpurun_out/${TAG}_bench_g8_$WL.err
  echo "bench N=8 $WL rc=$?"; cut -c1-160 gpurun_out/${TAG}_bench_g8_$WL.json; grep -E "rank 0\]|Error|error" gpurun_out/${TAG}_bench_g8_$WL.err | tail -3
done
echo "== timeline 8 GPUs"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 scripts/timeline.py --out gpurun_out/${TAG}_timeline_8gpu > /dev/null 2> gpurun_out/${TAG}_timeline_8gpu.err; echo "rc=$?"; head -24 gpurun_out/${TAG}_timeline_8gpu.txt | cut -c1-150
