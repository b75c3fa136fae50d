#!/usr/bin/env bash
# One gpurun call: GPU parity tests, smoke, bench (both arms), ncu launch list and one full capture of the hot kernels.
# Usage (from the repo root on the GPU box):  bash scripts/gpu_round.sh [tag]
set -u
TAG="${1:-r01}"
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/${TAG}_gpu.txt 2>&1
echo "== pytest -m gpu" ; timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/${TAG}_pytest_gpu.log
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $OUT/${TAG}_smoke.log
echo "== bench ours" ; timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; cat $OUT/${TAG}_bench.json; tail -5 $OUT/${TAG}_bench.err
echo "== bench reference" ; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err; echo "ref rc=$?"; cat $OUT/${TAG}_bench_ref.json
echo "== ncu launch list" ; timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 > $OUT/${TAG}_ncu_bench.log 2>&1; echo "ncu list rc=$?"
echo "== ncu full (msda + warp)" ; timeout 900 ncu --set full --clock-control none --import-source on -k regex:'msda_|warp_|transpose_|im2col|layernorm' -c 24 -f -o $OUT/${TAG}_prof python scripts/prof_kernels.py > $OUT/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la $OUT
if [ -n "${EXTRAS:-}" ]; then
echo "== sweep (configs 3/4)" ; timeout 900 python scripts/sweep.py --out $OUT/${TAG}_sweep.jsonl > $OUT/${TAG}_sweep.log 2>&1; echo "sweep rc=$?"; tail -3 $OUT/${TAG}_sweep.log | cut -c1-300
echo "== gemm modes" ; timeout 300 python scripts/exp_gemm.py > $OUT/${TAG}_gemm.jsonl 2> $OUT/${TAG}_gemm.err; echo "gemm rc=$?"; cut -c1-220 $OUT/${TAG}_gemm.jsonl | head -30; tail -3 $OUT/${TAG}_gemm.err
echo "== bwd diag" ; timeout 300 python scripts/diag_bwd.py > $OUT/${TAG}_diag_bwd.jsonl 2> $OUT/${TAG}_diag_bwd.err; echo "diag rc=$?"; cat $OUT/${TAG}_diag_bwd.jsonl; tail -3 $OUT/${TAG}_diag_bwd.err
echo "== bench strict torch gemm" ; MVDETR_B200_GEMM=torch timeout 600 python bench.py --steps 30 --warmup 5 > $OUT/${TAG}_bench_torchgemm.json 2> $OUT/${TAG}_bench_torchgemm.err; echo "rc=$?"; cut -c1-300 $OUT/${TAG}_bench_torchgemm.json
fi
