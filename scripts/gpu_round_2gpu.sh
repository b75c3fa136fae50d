mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR scripts/check_sharded.py > gpurun_out/r01c_sharded2.log 2>&1; echo "check rc=$?"; grep -E "wildtrack|multiviewx|SHARDED|Error|error" gpurun_out/r01c_sharded2.log | tail -12
timeout 300 $TR bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/r01c_bench_g2.json 2> gpurun_out/r01c_bench_g2.err; echo "bench2 rc=$?"; cut -c1-700 gpurun_out/r01c_bench_g2.json; tail -3 gpurun_out/r01c_bench_g2.err
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/r01c_bench_g1.json 2> gpurun_out/r01c_bench_g1.err; echo "bench1 rc=$?"; cut -c1-400 gpurun_out/r01c_bench_g1.json
