#!/usr/bin/env bash
# Round-2 GPU call f (1 GPU): bf16x3 GEMM v2 (two accumulators, smem bias, packed conversions): accuracy + speed + frame.
set -u
TAG="${1:-r02f}"
OUT=gpurun_out
mkdir -p $OUT
echo "== gemm tests"; timeout 400 python -m pytest tests/test_gemm_gpu.py -m gpu -q --timeout 100 2>&1 | tail -8
echo "== gemm bench"; timeout 300 python scripts/bench_gemm.py > $OUT/${TAG}_gemm.jsonl 2> $OUT/${TAG}_gemm.err; echo "rc=$?"; python - <<PY
import json
for l in open('gpurun_out/${TAG}_gemm.jsonl'):
    d=json.loads(l); print(d['name'], d['rows'],d['K'],d['N'],'floor',round(d['hbm_floor_us'],1), {k:(round(v,1) if k.endswith('_us') else float('%.2g'%v)) for k,v in d.items() if k.endswith('_us') and k!='hbm_floor_us' or k.endswith('_err')})
PY
tail -3 $OUT/${TAG}_gemm.err
echo "== bench ours (bf16x3)" ; MVDETR_B200_GEMM=bf16x3 timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench_bf16x3.json 2> $OUT/${TAG}_bench_bf16x3.err; echo "bench rc=$?"; cut -c1-200 $OUT/${TAG}_bench_bf16x3.json; tail -3 $OUT/${TAG}_bench_bf16x3.err
echo "== all gpu tests with bf16x3"; MVDETR_B200_GEMM=bf16x3 timeout 1200 python -m pytest tests -m gpu -q --timeout 600 --deselect tests/test_msda_gpu.py::test_reference_gradcheck_contract > $OUT/${TAG}_pytest_bf16x3.log 2>&1; echo "rc=$?"; tail -8 $OUT/${TAG}_pytest_bf16x3.log
echo "== timeline bf16x3"; MVDETR_B200_GEMM=bf16x3 timeout 300 python scripts/timeline.py --out $OUT/${TAG}_timeline_bf16x3 > /dev/null 2> $OUT/${TAG}_timeline.err; echo "rc=$?"; head -14 $OUT/${TAG}_timeline_bf16x3.txt | cut -c1-150
echo "== ncu full: gemm"; MVDETR_B200_GEMM=bf16x3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"linear_bf16x3" -c 3 -o $OUT/${TAG}_prof -f python scripts/prof_kernels.py > $OUT/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -2 $OUT/${TAG}_ncu_full.log
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/${TAG}_bench*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f,'unparsable',e); continue
    print(f, 'value',round(d.get('value',0),2),'ms',round(d.get('ms_per_step',0),3),'e2e',round(d.get('e2e',{}).get('value',0),2), 'gemm', d['config']['gemm'][:40])
    if 'ref_cuda_frame' in d: print('    ref_cuda_frame diff', d['ref_cuda_frame'].get('max_abs_diff_vs_ours'), d['ref_cuda_frame'].get('kernels_only'))
PY
