#!/usr/bin/env bash
# Round-2 GPU call s (1 GPU): state after the GEMM / MSDA work of r02m-r02r: whole suite, bench (3 workloads + CPU arm), ncu.
set -u
TAG="${1:-r02s}"
OUT=gpurun_out
mkdir -p $OUT
echo "== quick check"; timeout -s KILL 150 python scripts/ts_check.py 2>&1 | tail -3; RC=${PIPESTATUS[0]}; echo "quick rc=$RC"; [ "$RC" != "0" ] && exit 1
echo "== gemm bench"; timeout 300 python scripts/bench_gemm.py > $OUT/${TAG}_gemm.jsonl 2> $OUT/${TAG}_gemm.err; echo "rc=$?"; python - <<PY
import json
for l in open('gpurun_out/${TAG}_gemm.jsonl'):
    d=json.loads(l); print(d['name'], d['rows'],d['K'],d['N'],'floor',round(d['hbm_floor_us'],1), {k:(round(v,1) if k.endswith('_us') else float('%.2g'%v)) for k,v in d.items() if (k.endswith('_us') and k!='hbm_floor_us' or k.endswith('_err')) and not k.startswith('tf32') and not k.endswith('_err')})
PY
echo "== pytest -m gpu (all)"; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 --durations=6 > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 $OUT/${TAG}_pytest_gpu.log
echo "== bench ours" ; timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; cut -c1-300 $OUT/${TAG}_bench.json; tail -3 $OUT/${TAG}_bench.err
echo "== bench reference arm" ; timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err; echo "rc=$?"; cut -c1-200 $OUT/${TAG}_bench_ref.json
echo "== bench multiviewx"; timeout 600 python bench.py --workload multiviewx --steps 20 --warmup 5 > $OUT/${TAG}_bench_multiviewx.json 2> $OUT/${TAG}_bench_multiviewx.err; echo "rc=$?"; cut -c1-200 $OUT/${TAG}_bench_multiviewx.json
echo "== bench stress4k"; timeout 900 python bench.py --workload stress4k --steps 10 --warmup 3 > $OUT/${TAG}_bench_stress4k.json 2> $OUT/${TAG}_bench_stress4k.err; echo "rc=$?"; cut -c1-200 $OUT/${TAG}_bench_stress4k.json; tail -2 $OUT/${TAG}_bench_stress4k.err
echo "== timeline"; timeout 300 python scripts/timeline.py --out $OUT/${TAG}_timeline > /dev/null 2> $OUT/${TAG}_timeline.err; echo "rc=$?"; head -14 $OUT/${TAG}_timeline.txt | cut -c1-150
echo "== ncu launch list" ; timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 > $OUT/${TAG}_ncu_bench.log 2>&1; echo "ncu list rc=$?"
echo "== ncu full: every hot-path kernel"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:"msda_|warp_|linear_split|add_layernorm|upsample_im2col" -c 24 -o $OUT/${TAG}_prof -f python scripts/prof_kernels.py > $OUT/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -2 $OUT/${TAG}_ncu_full.log
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/${TAG}_bench*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f,'unparsable',e); continue
    print(f, 'value',round(d.get('value',0),2),'ms',round(d.get('ms_per_step',0),3),'e2e',round(d.get('e2e',{}).get('value',0),2), 'launches', d.get('gpu_launches'), 'roofline', round(d.get('roofline',{}).get('frac',0),3))
    if 'ref_cuda_frame' in d: print('    ref_cuda_frame', {k:v for k,v in d['ref_cuda_frame'].items() if k!='what'})
    if 'cpu_baseline' in d: print('    cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'])
PY
