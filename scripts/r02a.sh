#!/usr/bin/env bash
# Round-2 first GPU call: probes + full GPU test suite (new TMA warp, full-size parity) + the opt-in view-grid backward
# + bench (ours, reference arm) with A/B switches.
set -u
TAG="${1:-r02a}"
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tee $OUT/${TAG}_smi.txt
echo "== lds probe"; timeout 120 ./build/lds_probe | tee $OUT/${TAG}_lds_probe.txt
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 --durations=15 > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -32 $OUT/${TAG}_pytest_gpu.log
echo "== pytest view-grid backward (opt-in)"; MVDETR_B200_BWD_VIEWGRID=1 timeout 600 python -m pytest tests/test_msda_gpu.py tests/test_fullsize_gpu.py -m gpu -q --timeout 300 -k "viewgrid or vs_reference_cuda_op or gradcheck" > $OUT/${TAG}_pytest_bwdvg.log 2>&1; echo "rc=$?"; tail -15 $OUT/${TAG}_pytest_bwdvg.log
echo "== bench ours" ; timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; cut -c1-700 $OUT/${TAG}_bench.json; tail -5 $OUT/${TAG}_bench.err
echo "== bench reference arm" ; timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err; echo "rc=$?"; cut -c1-400 $OUT/${TAG}_bench_ref.json
echo "== bench, view-grid backward + old warp path (A/B)" ; MVDETR_B200_BWD_VIEWGRID=1 MVDETR_B200_WARP_TMA=0 timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench_ab.json 2> $OUT/${TAG}_bench_ab.err; echo "rc=$?"; cut -c1-300 $OUT/${TAG}_bench_ab.json
echo "== bench, GEMM autotune (A/B)" ; MVD_GEMM_AUTOTUNE=1 timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench_autotune.json 2> $OUT/${TAG}_bench_autotune.err; echo "rc=$?"; cut -c1-300 $OUT/${TAG}_bench_autotune.json
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02a_bench*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f,'unparsable',e); continue
    print(f, 'value',round(d.get('value',0),2),'ms',round(d.get('ms_per_step',0),3),'e2e',round(d.get('e2e',{}).get('value',0),2))
    for k,v in (d.get('kernels') or {}).items():
        if isinstance(v,dict): print('   ',k, {a:(round(b,1) if isinstance(b,float) else b) for a,b in v.items() if a in('us','GBps','kernel','launches')})
        else: print('   ',k,v)
    if 'ref_cuda_frame' in d: print('    ref_cuda_frame', d['ref_cuda_frame'])
    if 'cpu_baseline' in d: print('    cpu', d['cpu_baseline'])
PY
