#!/usr/bin/env bash
# Short GPU call: smoke, bench (ours), launch list of the timed steps, then the GPU tests with per-test durations.
set -u
TAG="${1:-q}"
OUT=gpurun_out
mkdir -p $OUT
echo "== smoke" ; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench ours" ; timeout 400 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; cut -c1-600 $OUT/${TAG}_bench.json; tail -4 $OUT/${TAG}_bench.err
echo "== bench cudnn convs" ; MVDETR_B200_CONV=cudnn timeout 300 python bench.py --steps 30 --warmup 5 > $OUT/${TAG}_bench_cudnnconv.json 2> /dev/null; echo "rc=$?"; cut -c1-260 $OUT/${TAG}_bench_cudnnconv.json
echo "== ncu launch list" ; timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 > $OUT/${TAG}_ncu_bench.log 2>&1; echo "ncu list rc=$?"
echo "== pytest -m gpu" ; timeout ${PYTEST_TIMEOUT:-420} python -m pytest tests -m gpu -q --durations=12 > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 $OUT/${TAG}_pytest_gpu.log
