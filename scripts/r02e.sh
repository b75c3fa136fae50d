#!/usr/bin/env bash
# Round-2 GPU call e (1 GPU): persistent bf16x3 tcgen05 GEMM (correctness under short timeouts first), warp v2b, A/B benches.
set -u
TAG="${1:-r02e}"
OUT=gpurun_out
mkdir -p $OUT
echo "== gemm tests"; timeout 400 python -m pytest tests/test_gemm_gpu.py -m gpu -q --timeout 100 -k "own_tensor or split_cache" 2>&1 | tail -15
echo "== gemm bench"; timeout 300 python scripts/bench_gemm.py > $OUT/${TAG}_gemm.jsonl 2> $OUT/${TAG}_gemm.err; echo "rc=$?"; python - <<'PY'
import json
for l in open('gpurun_out/r02e_gemm.jsonl'):
    d=json.loads(l); print(d['name'], d['rows'],d['K'],d['N'],'floor',round(d['hbm_floor_us'],1), {k:(round(v,1) if k.endswith('_us') else float('%.2g'%v)) for k,v in d.items() if k.endswith('_us') and k!='hbm_floor_us' or k.endswith('_err')})
PY
tail -3 $OUT/${TAG}_gemm.err
echo "== warp tests"; timeout 600 python -m pytest tests/test_warp_gpu.py -m gpu -q --timeout 300 2>&1 | tail -4
echo "== sanitizer: gemm bf16x3"; timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gemm_gpu.py -m gpu -q -x --timeout 500 -k "own_tensor and bf16x3 and (4099 or 130 or 641)" > $OUT/${TAG}_sanitizer.log 2>&1; echo "rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Illegal|Invalid|at mvd" $OUT/${TAG}_sanitizer.log | head
for MODE in bf16x3 tf32x3; do
echo "== bench ours ($MODE)" ; MVDETR_B200_GEMM=$MODE timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench_$MODE.json 2> $OUT/${TAG}_bench_$MODE.err; echo "bench rc=$?"; cut -c1-200 $OUT/${TAG}_bench_$MODE.json; tail -3 $OUT/${TAG}_bench_$MODE.err
done
echo "== all gpu tests with bf16x3"; MVDETR_B200_GEMM=bf16x3 timeout 1200 python -m pytest tests -m gpu -q --timeout 600 --deselect tests/test_msda_gpu.py::test_reference_gradcheck_contract > $OUT/${TAG}_pytest_bf16x3.log 2>&1; echo "rc=$?"; tail -8 $OUT/${TAG}_pytest_bf16x3.log
echo "== timeline bf16x3"; MVDETR_B200_GEMM=bf16x3 timeout 300 python scripts/timeline.py --out $OUT/${TAG}_timeline_bf16x3 > /dev/null 2> $OUT/${TAG}_timeline.err; echo "rc=$?"; head -16 $OUT/${TAG}_timeline_bf16x3.txt | cut -c1-150
echo "== ncu full: warp v2 + gemm"; MVDETR_B200_GEMM=bf16x3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"warp_tma_cl|linear_bf16x3" -c 5 -o $OUT/${TAG}_prof -f python scripts/prof_kernels.py > $OUT/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -2 $OUT/${TAG}_ncu_full.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02e_bench*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f,'unparsable',e); continue
    print(f, 'value',round(d.get('value',0),2),'ms',round(d.get('ms_per_step',0),3),'e2e',round(d.get('e2e',{}).get('value',0),2), 'gemm', d['config']['gemm'][:40])
    for k,v in (d.get('kernels') or {}).items():
        if isinstance(v,dict) and k.startswith('warp'): print('   ',k, {a:(round(b,1) if isinstance(b,float) else b) for a,b in v.items() if a in('us','GBps','kernel','launches')})
    if 'ref_cuda_frame' in d: print('    ref_cuda_frame diff', d['ref_cuda_frame'].get('max_abs_diff_vs_ours'), d['ref_cuda_frame'].get('kernels_only'))
PY
