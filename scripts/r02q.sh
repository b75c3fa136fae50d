#!/usr/bin/env bash
# Round-2 GPU call q (1 GPU): converged issue warps (elect.sync) in the GEMM, warp and MSDA kernels; zero-pad bank fix.
set -u
TAG="${1:-r02q}"
OUT=gpurun_out
mkdir -p $OUT
echo "== quick check"; timeout -s KILL 150 python scripts/ts_check.py 2>&1 | tail -14; RC=${PIPESTATUS[0]}; echo "quick rc=$RC"
if [ "$RC" != "0" ]; then
  echo "== sanitizer"; timeout -s KILL 300 compute-sanitizer --tool memcheck python scripts/ts_check.py > $OUT/${TAG}_sanitizer.log 2>&1
  echo "sanitizer rc=$?"; grep -E "err|ERROR SUMMARY|Invalid|Error|at 0x|by thread" $OUT/${TAG}_sanitizer.log | head -30
  exit 1
fi
echo "== gemm + msda + fullsize tests"; timeout -s KILL 900 python -m pytest tests/test_gemm_gpu.py tests/test_msda_gpu.py tests/test_fullsize_gpu.py tests/test_world_feat_gpu.py tests/test_warp_gpu.py -m gpu -q --timeout 300 -x 2>&1 | tail -6
echo "== gemm bench"; timeout -s KILL 400 python scripts/bench_gemm.py > $OUT/${TAG}_gemm.jsonl 2> $OUT/${TAG}_gemm.err; echo "rc=$?"; python - <<PY
import json
for l in open('gpurun_out/${TAG}_gemm.jsonl'):
    d=json.loads(l); print(d['name'], d['rows'],d['K'],d['N'],'floor',round(d['hbm_floor_us'],1), {k:(round(v,1) if k.endswith('_us') else float('%.2g'%v)) for k,v in d.items() if (k.endswith('_us') and k!='hbm_floor_us') and not k.startswith('tf32')})
PY
echo "== bench ours" ; timeout -s KILL 600 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; cut -c1-200 $OUT/${TAG}_bench.json; tail -3 $OUT/${TAG}_bench.err
echo "== bench stress4k"; timeout -s KILL 900 python bench.py --workload stress4k --steps 10 --warmup 3 > $OUT/${TAG}_bench_stress4k.json 2> $OUT/${TAG}_bench_stress4k.err; echo "rc=$?"; cut -c1-200 $OUT/${TAG}_bench_stress4k.json; tail -2 $OUT/${TAG}_bench_stress4k.err
echo "== timeline"; timeout -s KILL 300 python scripts/timeline.py --out $OUT/${TAG}_timeline > /dev/null 2> $OUT/${TAG}_timeline.err; echo "rc=$?"; head -14 $OUT/${TAG}_timeline.txt | cut -c1-150
echo "== ncu full: gemm big-K shapes"; timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"linear_split" -c 8 -o $OUT/${TAG}_prof_gemm -f python scripts/prof_gemm.py > $OUT/${TAG}_ncu_gemm.log 2>&1; echo "ncu rc=$?"; tail -2 $OUT/${TAG}_ncu_gemm.log
echo "== ncu full: msda + warp"; timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"msda_vg|warp_tma_cl" -c 5 -o $OUT/${TAG}_prof -f python scripts/prof_kernels.py > $OUT/${TAG}_ncu_full.log 2>&1; echo "ncu rc=$?"; tail -2 $OUT/${TAG}_ncu_full.log
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/${TAG}_bench*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f,'unparsable',e); continue
    print(f, 'value',round(d.get('value',0),2),'ms',round(d.get('ms_per_step',0),3),'e2e',round(d.get('e2e',{}).get('value',0),2), 'launches', d.get('gpu_launches'), 'roofline', round(d.get('roofline',{}).get('frac',0),3), 'msda us', round(d.get('roofline',{}).get('us_per_launch',0),1), 'clocks', d.get('clocks'))
    if 'ref_cuda_frame' in d: print('    ref_cuda_frame diff', d['ref_cuda_frame'].get('max_abs_diff_vs_ours'))
PY
