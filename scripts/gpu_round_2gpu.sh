#!/usr/bin/env bash
# 2-GPU box: view-sharded parity check, 2-rank bench, 1-rank bench (same box, back to back).
mkdir -p gpurun_out
TAG="${1:-r01c}"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
[ -n "$SKIP" ] || timeout 300 $TR scripts/check_sharded.py > gpurun_out/${TAG}_sharded2.log 2>&1; echo "check rc=$?"; grep -E "wildtrack|multiviewx|one_view|SHARDED|Error|error" gpurun_out/${TAG}_sharded2.log | tail -12
MVD_BENCH_TRACE=100 timeout 200 $TR bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/${TAG}_bench_g2.json 2> gpurun_out/${TAG}_bench_g2.err; echo "bench2 rc=$?"; cut -c1-900 gpurun_out/${TAG}_bench_g2.json; grep -v Warning gpurun_out/${TAG}_bench_g2.err | grep -E 'rank|File|line' | tail -60
[ -n "$SKIP" ] || timeout 300 python bench.py --steps 30 --warmup 5 > gpurun_out/${TAG}_bench_g1.json 2> gpurun_out/${TAG}_bench_g1.err; echo "bench1 rc=$?"; cut -c1-400 gpurun_out/${TAG}_bench_g1.json
