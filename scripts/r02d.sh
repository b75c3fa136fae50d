#!/usr/bin/env bash
# Round-2 GPU call d (1 GPU): tcgen05 3xTF32 GEMM (correctness first, under a short timeout), warp kernel v2, A/B benches.
set -u
TAG="${1:-r02d}"
OUT=gpurun_out
mkdir -p $OUT
echo "== tf32x3 gemm tests"; timeout 300 python -m pytest tests/test_gemm_gpu.py -m gpu -q -x --timeout 120 -k tf32x3 2>&1 | tail -15
echo "== gemm bench"; timeout 300 python scripts/bench_gemm.py > $OUT/${TAG}_gemm.jsonl 2> $OUT/${TAG}_gemm.err; echo "rc=$?"; cat $OUT/${TAG}_gemm.jsonl | cut -c1-400; tail -3 $OUT/${TAG}_gemm.err
echo "== warp tests (kernel v2)"; timeout 600 python -m pytest tests/test_warp_gpu.py -m gpu -q --timeout 300 2>&1 | tail -8
echo "== sanitizer: warp v2 + gemm"; timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_warp_gpu.py tests/test_gemm_gpu.py -m gpu -q -x --timeout 800 -k "(tma_warp_kernel_every_mode and (128-24 or 64-128 or 96-20)) or (tf32x3 and (4099 or 130))" > $OUT/${TAG}_sanitizer.log 2>&1; echo "rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Illegal|Invalid|at mvd" $OUT/${TAG}_sanitizer.log | head
echo "== bench ours (bf16x9 GEMMs)" ; timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; cut -c1-300 $OUT/${TAG}_bench.json; tail -3 $OUT/${TAG}_bench.err
echo "== bench ours (tf32x3 GEMMs)" ; MVDETR_B200_GEMM=tf32x3 timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench_tf32x3.json 2> $OUT/${TAG}_bench_tf32x3.err; echo "bench rc=$?"; cut -c1-300 $OUT/${TAG}_bench_tf32x3.json; tail -3 $OUT/${TAG}_bench_tf32x3.err
echo "== full-size parity with tf32x3"; MVDETR_B200_GEMM=tf32x3 timeout 600 python -m pytest tests/test_fullsize_gpu.py tests/test_world_feat_gpu.py -m gpu -q --timeout 300 -k "frame_runner or world_feat or fusion" 2>&1 | tail -6
echo "== timeline tf32x3"; MVDETR_B200_GEMM=tf32x3 timeout 300 python scripts/timeline.py --out $OUT/${TAG}_timeline_tf32x3 > /dev/null 2> $OUT/${TAG}_timeline.err; echo "rc=$?"; head -16 $OUT/${TAG}_timeline_tf32x3.txt | cut -c1-150
echo "== ncu full: warp v2 + gemm"; MVDETR_B200_GEMM=tf32x3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"warp_tma_cl|linear_tf32x3" -c 4 -o $OUT/${TAG}_prof -f python scripts/prof_kernels.py > $OUT/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -2 $OUT/${TAG}_ncu_full.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02d_bench*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f,'unparsable',e); continue
    print(f, 'value',round(d.get('value',0),2),'ms',round(d.get('ms_per_step',0),3),'e2e',round(d.get('e2e',{}).get('value',0),2), 'gemm', d['config']['gemm'][:40])
    for k,v in (d.get('kernels') or {}).items():
        if isinstance(v,dict) and k.startswith('warp'): print('   ',k, {a:(round(b,1) if isinstance(b,float) else b) for a,b in v.items() if a in('us','GBps','kernel','launches')})
    if 'ref_cuda_frame' in d: print('    ref_cuda_frame diff', d['ref_cuda_frame'].get('max_abs_diff_vs_ours'), d['ref_cuda_frame'].get('kernels_only'))
PY
