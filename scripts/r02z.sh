#!/usr/bin/env bash
# Round-2 GPU call z (1 GPU): last validation of the round: smoke, whole suite, bench (3 workloads + CPU arm), launch list.
# whole suite, bench (3 workloads + CPU arm), launch list, timeline, smoke.
set -u
TAG="${1:-r02z}"
OUT=gpurun_out
mkdir -p $OUT
echo "== smoke"; timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
echo "== pytest -m gpu (all)"; timeout -s KILL 1500 python -m pytest tests -m gpu -q --timeout 600 --durations=4 > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -9 $OUT/${TAG}_pytest_gpu.log
echo "== bench ours" ; timeout -s KILL 600 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; cut -c1-200 $OUT/${TAG}_bench.json; tail -2 $OUT/${TAG}_bench.err
echo "== bench reference arm" ; timeout -s KILL 600 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err; echo "rc=$?"; cut -c1-200 $OUT/${TAG}_bench_ref.json
echo "== bench multiviewx"; timeout -s KILL 600 python bench.py --workload multiviewx --steps 20 --warmup 5 > $OUT/${TAG}_bench_multiviewx.json 2> $OUT/${TAG}_bench_multiviewx.err; echo "rc=$?"; cut -c1-200 $OUT/${TAG}_bench_multiviewx.json
echo "== bench stress4k"; timeout -s KILL 900 python bench.py --workload stress4k --steps 10 --warmup 3 > $OUT/${TAG}_bench_stress4k.json 2> $OUT/${TAG}_bench_stress4k.err; echo "rc=$?"; cut -c1-200 $OUT/${TAG}_bench_stress4k.json; tail -2 $OUT/${TAG}_bench_stress4k.err
echo "== ncu launch list" ; timeout -s KILL 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 > $OUT/${TAG}_ncu_bench.log 2>&1; echo "ncu list rc=$?"
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/${TAG}_bench*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f,'unparsable',e); continue
    print(f, 'value',round(d.get('value',0),2),'ms',round(d.get('ms_per_step',0),3),'e2e',round(d.get('e2e',{}).get('value',0),2), 'launches', d.get('gpu_launches'), 'roofline', round(d.get('roofline',{}).get('frac',0),3))
    if 'ref_cuda_frame' in d: print('    ref_cuda_frame', d['ref_cuda_frame'].get('max_abs_diff_vs_ours'), d['ref_cuda_frame'].get('ours_over_ref_kernels_only'), d['ref_cuda_frame'].get('ours_over_ref_as_shipped'))
PY
