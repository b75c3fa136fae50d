#!/usr/bin/env bash
# Round-2 GPU call m (1 GPU): first run of the TS-form GEMM (x terms in tensor memory): memcheck on small shapes, accuracy +
# bit-identity tests, per-shape timing against the shared-memory kernel, frame bench with either.
set -u
TAG="${1:-r02m}"
OUT=gpurun_out
mkdir -p $OUT
cat > /tmp/ts_check.py <<'PY'
import torch
from mvdetr_b200 import ops
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
ok = True
for rows, K, N in [(300, 128, 448), (129, 64, 128), (1000, 512, 128), (257, 288, 256), (5, 8, 4), (75600, 128, 512)]:
    x = torch.randn(rows, K, generator=g).to(dev); w = (torch.randn(N, K, generator=g) / K ** 0.5).to(dev); b = torch.randn(N, generator=g).to(dev)
    a = ops.linear(x, w, b, mode="bf16x3ts"); s = ops.linear(x, w, b, mode="bf16x3ss")
    torch.cuda.synchronize()
    e = (a.double() - (x.double() @ w.double().t() + b.double())).abs().max().item()
    print(rows, K, N, "err", e, "equal_to_ss", torch.equal(a, s), "max diff to ss", (a - s).abs().max().item(), flush=True)
    ok = ok and e < 1e-4
raise SystemExit(0 if ok else 1)
PY
echo "== quick check (TS kernel)"; timeout -s KILL 150 python /tmp/ts_check.py 2>&1 | tail -12; RC=${PIPESTATUS[0]}; echo "quick rc=$RC"
if [ "$RC" != "0" ]; then
  echo "== sanitizer (TS kernel, small shapes)"; timeout -s KILL 300 compute-sanitizer --tool memcheck python /tmp/ts_check.py > $OUT/${TAG}_sanitizer_ts.log 2>&1
  echo "sanitizer rc=$?"; grep -E "err|ERROR SUMMARY|Invalid|Error|at 0x|by thread" $OUT/${TAG}_sanitizer_ts.log | head -30
  exit 1
fi
echo "== gemm tests"; timeout -s KILL 500 python -m pytest tests/test_gemm_gpu.py -m gpu -q --timeout 100 -x 2>&1 | tail -8
echo "== gemm bench"; timeout -s KILL 400 python scripts/bench_gemm.py > $OUT/${TAG}_gemm.jsonl 2> $OUT/${TAG}_gemm.err; echo "rc=$?"; python - <<PY
import json
for l in open('gpurun_out/${TAG}_gemm.jsonl'):
    d=json.loads(l); print(d['name'], d['rows'],d['K'],d['N'],'floor',round(d['hbm_floor_us'],1), {k:(round(v,1) if k.endswith('_us') else float('%.2g'%v)) for k,v in d.items() if (k.endswith('_us') and k!='hbm_floor_us' or k.endswith('_err') or k.endswith('_error')) and not k.startswith('tf32')})
PY
for TS in 1 0; do
echo "== bench ours TS=$TS" ; MVDETR_B200_GEMM_TS=$TS timeout -s KILL 600 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench_ts$TS.json 2> $OUT/${TAG}_bench_ts$TS.err; echo "bench rc=$?"; cut -c1-200 $OUT/${TAG}_bench_ts$TS.json; tail -3 $OUT/${TAG}_bench_ts$TS.err
done
echo "== fullsize parity with TS"; MVDETR_B200_GEMM_TS=1 timeout -s KILL 600 python -m pytest tests/test_fullsize_gpu.py tests/test_world_feat_gpu.py -m gpu -q --timeout 300 2>&1 | tail -4
echo "== timeline TS"; MVDETR_B200_GEMM_TS=1 timeout -s KILL 300 python scripts/timeline.py --out $OUT/${TAG}_timeline_ts > /dev/null 2> $OUT/${TAG}_timeline.err; echo "rc=$?"; head -14 $OUT/${TAG}_timeline_ts.txt | cut -c1-150
