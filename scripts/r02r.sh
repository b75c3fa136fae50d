#!/usr/bin/env bash
# Round-2 GPU call r (1 GPU): ring-depth sweep of the TS GEMM (x-tile ring vs weight-term ring), then the frame with the best.
set -u
TAG="${1:-r02r}"
OUT=gpurun_out
mkdir -p $OUT
echo "== quick check"; timeout -s KILL 150 python scripts/ts_check.py 2>&1 | tail -4; RC=${PIPESTATUS[0]}; echo "quick rc=$RC"
[ "$RC" != "0" ] && exit 1
: > $OUT/${TAG}_rings.jsonl
for R in default 8,3 7,4 6,5 5,6 4,7 3,8 4,4 3,3; do
  if [ "$R" = "default" ]; then unset MVD_GEMM_RINGS; else export MVD_GEMM_RINGS=$R; fi
  BENCH_GEMM_MODES=f16x2 timeout -s KILL 200 python scripts/bench_gemm.py >> $OUT/${TAG}_rings.jsonl 2>> $OUT/${TAG}_rings.err
done
unset MVD_GEMM_RINGS
for R in 7,3 5,4 4,5 3,6; do
  MVD_GEMM_RINGS=$R BENCH_GEMM_MODES=bf16x3ts timeout -s KILL 200 python scripts/bench_gemm.py >> $OUT/${TAG}_rings.jsonl 2>> $OUT/${TAG}_rings.err
done
python - <<PY
import json, collections
t=collections.OrderedDict()
for l in open('gpurun_out/${TAG}_rings.jsonl'):
    d=json.loads(l)
    for m in ('f16x2','bf16x3ts'):
        if m+'_us' in d: t.setdefault((m,d['rings']),{})[d['name']]=d[m+'_us']
names=None
for k,v in t.items():
    if names is None: names=list(v); print('mode rings', ' '.join(n[:9] for n in names), 'frame_sum')
    fs=3*(2*v['value/out_proj']+v['offsets']+v['logits']+v['linear1']+v['linear2'])+v['conv_down']+v['conv_up']+v['merge']
    print(k[0], k[1], ' '.join('%9.1f'%v[n] for n in names), '%.0f'%fs)
PY
echo "== bench ours (default rings)"; timeout -s KILL 600 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; cut -c1-200 $OUT/${TAG}_bench.json; tail -2 $OUT/${TAG}_bench.err
