#!/usr/bin/env bash
# Round-2 GPU call ab (1 GPU): error budget of the full-size frame (ours vs fp32 / fp64 CPU oracle) for both GEMM splits.
set -u
TAG="${1:-r02ab}"
mkdir -p gpurun_out
: > gpurun_out/${TAG}_frame_error.jsonl
for MODE in f16x2 bf16x3; do
  MVDETR_B200_GEMM=$MODE timeout -s KILL 400 python tests/frame_error.py >> gpurun_out/${TAG}_frame_error.jsonl 2>> gpurun_out/${TAG}_frame_error.err
done
cat gpurun_out/${TAG}_frame_error.jsonl; tail -3 gpurun_out/${TAG}_frame_error.err
