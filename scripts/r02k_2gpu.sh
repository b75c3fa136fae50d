#!/usr/bin/env bash
# quick 2-GPU validation of the sharded path before the 8-GPU call
mkdir -p gpurun_out
TAG="${1:-r02k}"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port"
echo "== sharded check, fused multicast gather"; timeout 400 $TR 29511 scripts/check_sharded.py > gpurun_out/${TAG}_sharded2_fused.log 2>&1; echo "check rc=$?"; grep -E "SHARDED|Error|error" gpurun_out/${TAG}_sharded2_fused.log | tail -5; grep -c "0.00e+00', '0.00e+00'" gpurun_out/${TAG}_sharded2_fused.log
echo "== bench 2 GPUs fused"; MVD_BENCH_TRACE=150 timeout 300 $TR 29513 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/${TAG}_bench_g2.json 2> gpurun_out/${TAG}_bench_g2.err; echo "bench2 rc=$?"; cut -c1-200 gpurun_out/${TAG}_bench_g2.json; grep -E 'rank 0\]|Error|error' gpurun_out/${TAG}_bench_g2.err | tail -5
echo "== timeline 2 GPUs"; timeout 300 $TR 29515 scripts/timeline.py --out gpurun_out/${TAG}_timeline_2gpu > /dev/null 2> gpurun_out/${TAG}_timeline_2gpu.err; echo "rc=$?"; head -16 gpurun_out/${TAG}_timeline_2gpu.txt | cut -c1-150
