#!/usr/bin/env bash
# Round-2 2-GPU call: view-sharded parity (NCCL and fused multicast gathers), 2-rank bench both ways, timeline.
mkdir -p gpurun_out
TAG="${1:-r02h}"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port"
echo "== sharded check, fused multicast gather"; timeout 400 $TR 29511 scripts/check_sharded.py > gpurun_out/${TAG}_sharded2_fused.log 2>&1; echo "check rc=$?"; grep -E "wildtrack|multiviewx|one_view|SHARDED|Error|error|mode" gpurun_out/${TAG}_sharded2_fused.log | tail -14
echo "== sharded check, NCCL gathers"; MVDETR_B200_FUSED_GATHER=0 timeout 400 $TR 29512 scripts/check_sharded.py > gpurun_out/${TAG}_sharded2_nccl.log 2>&1; echo "check rc=$?"; grep -E "wildtrack|multiviewx|one_view|SHARDED|Error|error" gpurun_out/${TAG}_sharded2_nccl.log | tail -14
echo "== bench 2 GPUs fused"; MVD_BENCH_TRACE=150 timeout 300 $TR 29513 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/${TAG}_bench_g2.json 2> gpurun_out/${TAG}_bench_g2.err; echo "bench2 rc=$?"; cut -c1-700 gpurun_out/${TAG}_bench_g2.json; grep -v Warning gpurun_out/${TAG}_bench_g2.err | grep -E 'rank 0|Error|error|File' | tail -12
echo "== bench 2 GPUs nccl"; MVDETR_B200_FUSED_GATHER=0 MVD_BENCH_TRACE=150 timeout 300 $TR 29514 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/${TAG}_bench_g2_nccl.json 2> gpurun_out/${TAG}_bench_g2_nccl.err; echo "bench2 rc=$?"; cut -c1-700 gpurun_out/${TAG}_bench_g2_nccl.json; grep -E 'rank 0|Error|error' gpurun_out/${TAG}_bench_g2_nccl.err | tail -5
echo "== timeline 2 GPUs"; timeout 300 $TR 29515 scripts/timeline.py --out gpurun_out/${TAG}_timeline_2gpu > /dev/null 2> gpurun_out/${TAG}_timeline_2gpu.err; echo "rc=$?"; head -24 gpurun_out/${TAG}_timeline_2gpu.txt | cut -c1-150
echo "== bench 1 GPU (same box)"; timeout 300 python bench.py --steps 30 --warmup 5 > gpurun_out/${TAG}_bench_g1.json 2> gpurun_out/${TAG}_bench_g1.err; echo "bench1 rc=$?"; cut -c1-300 gpurun_out/${TAG}_bench_g1.json
