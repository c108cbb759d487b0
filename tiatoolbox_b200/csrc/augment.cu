// augment.cu — StainAugmentor.augment on the device (tiatoolbox/tools/stainaugment.py:171-206).
//
// augment() scales / shifts the two stain concentrations of every (tissue) pixel by random (alpha_i, beta_i) and recomposes
// 255 * exp(-(C' . M)), clipped and truncated to uint8.  One elementwise kernel: reads 8 B (two float32 concentrations) + 1 B (mask),
// writes 3 B per pixel — HBM-bound.  The 2x3 stain matrix and the four scalars travel as kernel arguments.
#include <cuda_runtime.h>

#include <string>

#include "../../include/tiab200.h"
#include "ctx.h"

namespace tiab {

struct AugmentArgs {
  float m[6];        // stain matrix rows (H, E) x (R, G, B), pre-multiplied by log2(e) so that exp() is one ex2
  float alpha[2];
  float beta[2];
};

__global__ void __launch_bounds__(256) stain_augment_kernel(const float2* __restrict__ conc, const uint8_t* __restrict__ mask, long long npix,
                                                            AugmentArgs a, uint8_t* __restrict__ out) {
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += (long long)gridDim.x * blockDim.x) {
    float2 c = __ldg(conc + p);
    if (mask == nullptr || mask[p]) {   // augmented_concentrations[tissue_mask, i] *= alpha; += beta (stainaugment.py:189-196)
      c.x = __fadd_rn(__fmul_rn(c.x, a.alpha[0]), a.beta[0]);
      c.y = __fadd_rn(__fmul_rn(c.y, a.alpha[1]), a.beta[1]);
    }
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      float v = 255.0f * exp2f(-(c.x * a.m[ch] + c.y * a.m[3 + ch]));
      v = fminf(fmaxf(v, 0.0f), 255.0f);      // np.clip; NaN (0 * inf) -> fmaxf gives 0
      out[p * 3 + ch] = (uint8_t)v;            // astype(uint8): truncation
    }
  }
}

}  // namespace tiab

using namespace tiab;

extern "C" int tiab_stain_augment(tiab_handle_t h, const float* conc_dev, const uint8_t* mask_dev, long long npix, const double M[6],
                                  const double alpha[2], const double beta[2], uint8_t* out_dev, void* stream) {
  if (!h) return TIAB_E_INVALID;
  if (!conc_dev || npix <= 0 || !M || !alpha || !beta || !out_dev) {
    h->err = "tiab_stain_augment: invalid argument";
    return TIAB_E_INVALID;
  }
  TIAB_ON_DEVICE(h);
  AugmentArgs a;
  for (int i = 0; i < 6; ++i) a.m[i] = (float)(M[i] * 1.4426950408889634);
  for (int i = 0; i < 2; ++i) a.alpha[i] = (float)alpha[i], a.beta[i] = (float)beta[i];
  const long long blocks = (npix + 255) / 256;
  const int grid = (int)(blocks < (long long)h->sm_count * 16 ? blocks : (long long)h->sm_count * 16);
  stain_augment_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2*>(conc_dev), mask_dev, npix, a, out_dev);
  h->last_launches = 1;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    h->err = std::string("stain_augment_kernel: ") + cudaGetErrorString(e);
    return TIAB_E_CUDA;
  }
  return TIAB_OK;
}
