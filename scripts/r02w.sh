#!/usr/bin/env bash
# Round-2 GPU call w (1 GPU): band-by-band pair order in the MSDA backward: tests, timing of both orders, DRAM traffic.
set -u
TAG="${1:-r02w}"
OUT=gpurun_out
mkdir -p $OUT
echo "== msda tests"; timeout -s KILL 900 python -m pytest tests/test_msda_gpu.py tests/test_fullsize_gpu.py -m gpu -q --timeout 300 -x 2>&1 | tail -5
echo "== bench (backward timing in kernels.msda_bwd / msda_bwd_query_order)"; timeout -s KILL 600 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1])
print('value',round(d['value'],1),'ms',round(d['ms_per_step'],3))
for k in ('msda_bwd','msda_bwd_query_order','ref_cuda_msda_bwd','msda_fused_fwd'):
    print(k, d['kernels'].get(k))
PY
echo "== ncu: DRAM traffic of both orders"
for B in 1 0; do
MVDETR_B200_BWD_BANDED=$B timeout -s KILL 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:msda_bwd -c 1 --csv --log-file $OUT/${TAG}_ncu_bwd_banded$B.csv python scripts/prof_kernels.py > /dev/null 2>&1; echo "ncu rc=$?"; tail -6 $OUT/${TAG}_ncu_bwd_banded$B.csv | cut -d, -f5,13,15 | tr -d '"'
done
