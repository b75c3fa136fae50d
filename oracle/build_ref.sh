#!/usr/bin/env bash
# Builds the reference's MultiScaleDeformableAttention CUDA extension for sm_100a into oracle/_ref/ (git-ignored,
# but shipped to the GPU box). Needs /root/reference; a no-op message otherwise. See ref_msda_wrapper.cu.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${MVDETR_REFERENCE:-/root/reference}"
SRC="$REF/multiview_detector/models/ops/src"
OUT="$HERE/_ref"
if [ ! -d "$SRC" ]; then echo "reference sources not found at $SRC; skipping oracle/_ref"; exit 0; fi
mkdir -p "$OUT"
PY="${PYTHON:-python}"
TORCH_INC=$($PY -c "import torch.utils.cpp_extension as c; print(' '.join('-I'+p for p in c.include_paths()))")
TORCH_LIB=$($PY -c "import torch.utils.cpp_extension as c; print(c.library_paths()[0])")
PY_INC=$($PY -c "import sysconfig; print(sysconfig.get_paths()['include'])")
NVCC="${NVCC:-$(command -v nvcc || echo /usr/local/cuda/bin/nvcc)}"
"$NVCC" -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo --expt-relaxed-constexpr \
  -Xcompiler -fPIC -w -shared \
  -DWITH_CUDA -DTORCH_EXTENSION_NAME=MultiScaleDeformableAttention -DTORCH_API_INCLUDE_EXTENSION_H \
  -DCUDA_HAS_FP16=1 -D__CUDA_NO_HALF_OPERATORS__ -D__CUDA_NO_HALF_CONVERSIONS__ -D__CUDA_NO_HALF2_OPERATORS__ \
  -D_GLIBCXX_USE_CXX11_ABI=$($PY -c "import torch; print(int(torch._C._GLIBCXX_USE_CXX11_ABI))") \
  -I "$SRC" $TORCH_INC -I "$PY_INC" \
  "$HERE/ref_msda_wrapper.cu" -o "$OUT/MultiScaleDeformableAttention.so" \
  -L "$TORCH_LIB" -lc10 -ltorch -ltorch_cpu -ltorch_python -lc10_cuda -ltorch_cuda \
  -Xlinker -rpath -Xlinker "$TORCH_LIB"
echo "built $OUT/MultiScaleDeformableAttention.so"
