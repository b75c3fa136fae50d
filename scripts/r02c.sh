#!/usr/bin/env bash
# Round-2 GPU call c (1 GPU): TMA warp after the 16-byte coordinate fix, full suite, bench, launch list, timeline, ncu of the warp.
set -u
TAG="${1:-r02c}"
OUT=gpurun_out
mkdir -p $OUT
echo "== lds probe"; timeout 120 ./build/lds_probe | tee $OUT/${TAG}_lds_probe.txt
echo "== tma warp tests"; timeout 600 python -m pytest tests/test_warp_gpu.py tests/test_preprocess.py -m gpu -q --timeout 300 2>&1 | tail -8
echo "== sanitizer on one tma case"; timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_warp_gpu.py -m gpu -q -x --timeout 500 -k "tma_warp_kernel_every_mode and 128-24" > $OUT/${TAG}_sanitizer_warp.log 2>&1; echo "rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Illegal|Invalid" $OUT/${TAG}_sanitizer_warp.log | head
echo "== pytest -m gpu (all)"; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 --durations=8 > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -14 $OUT/${TAG}_pytest_gpu.log
echo "== bench ours" ; timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; cut -c1-400 $OUT/${TAG}_bench.json; tail -3 $OUT/${TAG}_bench.err
echo "== bench, GEMM autotune (A/B)" ; MVD_GEMM_AUTOTUNE=1 timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench_autotune.json 2> $OUT/${TAG}_bench_autotune.err; echo "rc=$?"; cut -c1-200 $OUT/${TAG}_bench_autotune.json
echo "== bench multiviewx"; timeout 600 python bench.py --workload multiviewx --steps 20 --warmup 5 > $OUT/${TAG}_bench_multiviewx.json 2> $OUT/${TAG}_bench_multiviewx.err; echo "rc=$?"; cut -c1-200 $OUT/${TAG}_bench_multiviewx.json
echo "== bench stress4k"; timeout 900 python bench.py --workload stress4k --steps 10 --warmup 3 > $OUT/${TAG}_bench_stress4k.json 2> $OUT/${TAG}_bench_stress4k.err; echo "rc=$?"; cut -c1-200 $OUT/${TAG}_bench_stress4k.json; tail -3 $OUT/${TAG}_bench_stress4k.err
echo "== timeline"; timeout 300 python scripts/timeline.py --out $OUT/${TAG}_timeline > /dev/null 2> $OUT/${TAG}_timeline.err; echo "rc=$?"; head -40 $OUT/${TAG}_timeline.txt
echo "== ncu launch list" ; timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 > $OUT/${TAG}_ncu_bench.log 2>&1; echo "ncu list rc=$?"
echo "== ncu full: warp + msda kernels"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:"warp_tma|msda_vg_kernel" -c 4 -o $OUT/${TAG}_prof -f python scripts/prof_kernels.py > $OUT/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -3 $OUT/${TAG}_ncu_full.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02c_bench*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f,'unparsable',e); continue
    print(f, 'value',round(d.get('value',0),2),'ms',round(d.get('ms_per_step',0),3),'e2e',round(d.get('e2e',{}).get('value',0),2), 'roofline frac', round(d['roofline']['frac'],3), 'warp', d['hot_path']['warp_us'], d['hot_path']['warp_frac'])
    for k,v in (d.get('kernels') or {}).items():
        if isinstance(v,dict): print('   ',k, {a:(round(b,1) if isinstance(b,float) else b) for a,b in v.items() if a in('us','GBps','kernel','launches')})
        else: print('   ',k,v)
    if 'ref_cuda_frame' in d: print('    ref_cuda_frame', d['ref_cuda_frame'])
PY
