// View-grid (MVDeTr encoder layout) forward: TMA-staged value windows in shared memory.
// Placeholder until the tiled kernel lands: reports MVD_ERR_UNSUPPORTED so callers use mvd_msda_fwd_f32.
#include "common.cuh"

extern "C" int mvd_msda_fwd_viewgrid_f32(const float* value, const float* loc, const float* attn, int B, int H,
                                         int W, int M, int D, int L, int R, int P, float* out, void* stream) {
  (void)value; (void)loc; (void)attn; (void)B; (void)H; (void)W; (void)M; (void)D; (void)L; (void)R; (void)P;
  (void)out; (void)stream;
  return MVD_ERR_UNSUPPORTED;
}
