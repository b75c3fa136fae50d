#!/usr/bin/env bash
# Round-2 GPU call u (1 GPU): implicit convs + query from the conv epilogue + pinned upsample arithmetic: suite, bench, timeline.
set -u
TAG="${1:-r02u}"
OUT=gpurun_out
mkdir -p $OUT
echo "== conv check"; timeout -s KILL 200 python scripts/conv_check.py 2>&1 | tail -8; RC=${PIPESTATUS[0]}; echo "conv rc=$RC"
if [ "$RC" != "0" ]; then
  echo "== sanitizer"; timeout -s KILL 400 compute-sanitizer --tool memcheck python scripts/conv_check.py > $OUT/${TAG}_sanitizer.log 2>&1
  echo "sanitizer rc=$?"; grep -E "err|ERROR SUMMARY|Invalid|Error|at 0x|by thread" $OUT/${TAG}_sanitizer.log | head -30
  echo "== bench with the im2col route"; MVDETR_B200_CONV3X3=im2col timeout -s KILL 600 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench_im2col.json 2> $OUT/${TAG}_bench_im2col.err; cut -c1-200 $OUT/${TAG}_bench_im2col.json
  exit 1
fi
echo "== pytest -m gpu (all)"; timeout -s KILL 1500 python -m pytest tests -m gpu -q --timeout 600 --durations=4 > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -9 $OUT/${TAG}_pytest_gpu.log
for MODE in implicit; do
echo "== bench ours CONV3X3=$MODE" ; MVDETR_B200_CONV3X3=$MODE timeout -s KILL 600 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench_$MODE.json 2> $OUT/${TAG}_bench_$MODE.err; echo "bench rc=$?"; cut -c1-200 $OUT/${TAG}_bench_$MODE.json; tail -2 $OUT/${TAG}_bench_$MODE.err
done
echo "== timeline"; timeout -s KILL 300 python scripts/timeline.py --out $OUT/${TAG}_timeline > /dev/null 2> $OUT/${TAG}_timeline.err; echo "rc=$?"; head -14 $OUT/${TAG}_timeline.txt | cut -c1-150
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/${TAG}_bench*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f,'unparsable',e); continue
    print(f, 'value',round(d.get('value',0),2),'ms',round(d.get('ms_per_step',0),3),'e2e',round(d.get('e2e',{}).get('value',0),2), 'launches', d.get('gpu_launches'), 'roofline', round(d.get('roofline',{}).get('frac',0),3), 'warp', d.get('hot_path',{}).get('warp_kernel'), round(d.get('hot_path',{}).get('warp_us',0),1))
    if 'ref_cuda_frame' in d: print('    ref_cuda_frame diff', d['ref_cuda_frame'].get('max_abs_diff_vs_ours'))
PY
