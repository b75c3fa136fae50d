// oracle/ref_msda_wrapper.cu -- TEST INFRASTRUCTURE. Compiles the REFERENCE's own CUDA op for sm_100a, from its
// sources where they lie under /root/reference (nothing is copied into this repo), into
// oracle/_ref/MultiScaleDeformableAttention.so. Used as the on-GPU parity oracle and as the
// "reference's own ops/ CUDA path" comparator in bench.py.
//
// The reference passes `value.type()` (DeprecatedTypeProperties) to AT_DISPATCH_FLOATING_TYPES
// (multiview_detector/models/ops/src/cuda/ms_deform_attn_cuda.cu:64,134), which torch >= 2.x rejects. Instead of
// patching the reference file, the dispatch macro is re-defined here (after ATen's headers, whose include guards
// keep this definition in force) to accept that argument type. No reference line is modified.
#include <ATen/ATen.h>
#include <ATen/Dispatch.h>
#include <torch/extension.h>

#undef AT_DISPATCH_FLOATING_TYPES
#define AT_DISPATCH_FLOATING_TYPES(TYPE, NAME, ...) \
  AT_DISPATCH_SWITCH((TYPE).scalarType(), NAME, AT_DISPATCH_CASE_FLOATING_TYPES(__VA_ARGS__))

#include "cuda/ms_deform_attn_cuda.cu"   // found through -I $(REF)/multiview_detector/models/ops/src
#include "cpu/ms_deform_attn_cpu.cpp"
#include "vision.cpp"
