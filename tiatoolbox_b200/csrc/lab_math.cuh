// lab_math.cuh — OpenCV's 8-bit fixed-point RGB <-> CIE Lab (D65, sRGB gamma) as integer arithmetic, shared by the device kernels
// and the host-side test hooks (tiab_dbg_lab_convert), so that the exact code that runs on the GPU is compared with cv2 on the CPU.
//
// Reference call sites: tiatoolbox/tools/stainnorm.py:305 (cv2.cvtColor(img, COLOR_RGB2LAB) in ReinhardNormalizer.lab_split) and
// :334 (cv2.cvtColor(img, COLOR_LAB2RGB) in merge_back); utils/misc.py:281 (the L channel of the tissue mask, csrc/stain_math.cuh).
// OpenCV (the reference's pinned dependency, opencv-python >= 4.8) converts uint8 images with table-driven fixed-point code
// (imgproc/color_lab.cpp: RGB2Lab_b, Lab2RGBinteger); restated here from its published algorithm and pinned by an exhaustive
// comparison with cv2 4.13 over all 2^24 inputs in both directions (tests/test_oracle_lab.py): 0 mismatches.
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define TIAB_LAB_HD __host__ __device__ __forceinline__
#else
#define TIAB_LAB_HD inline
#endif

namespace tiab {

// Pointers to the four tables (csrc/lab_tables.h); on the device they point into shared memory.
struct LabTables {
  const uint16_t* gamma;      // [256]   sRGB gamma, scale 2040
  const uint16_t* cbrt;       // [2041]  f(t) = cbrt / linear segment, scale 32768
  const uint16_t* yf;         // [512]   per 8-bit L: (Y, fY), scale 16384
  const uint8_t* inv_gamma;   // [4096]  inverse sRGB gamma -> 0..255
};

TIAB_LAB_HD int lab_sat_u8(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

// cv2.cvtColor(COLOR_RGB2LAB) for one uint8 pixel.  XYZ rows are the sRGB/D65 matrix divided by the white point, scale 4096.
TIAB_LAB_HD void rgb_to_lab_u8(const LabTables& t, int r8, int g8, int b8, int& L, int& a, int& b) {
  const int R = t.gamma[r8], G = t.gamma[g8], B = t.gamma[b8];
  const int fX = t.cbrt[(1777 * R + 1541 * G + 778 * B + 2048) >> 12];
  const int fY = t.cbrt[(871 * R + 2929 * G + 296 * B + 2048) >> 12];
  const int fZ = t.cbrt[(73 * R + 448 * G + 3575 * B + 2048) >> 12];
  L = lab_sat_u8((296 * fY - 1336934 + 16384) >> 15);                 // 116 f(Y) - 16, scaled by 255 / 100
  a = lab_sat_u8((500 * (fX - fY) + 128 * 32768 + 16384) >> 15);      // arithmetic shift: floor for negative values
  b = lab_sat_u8((200 * (fY - fZ) + 128 * 32768 + 16384) >> 15);
}

// f^-1 of the inverse path in units of 1/16384: linear below 6/29, cubic above (OpenCV's abToXZ_b table, computed instead of stored)
TIAB_LAB_HD int lab_finv_q14(int i) {
  if (i <= 3390) return (i * 108) / 841 - 16384 * 16 / 116 * 108 / 841;          // C division: truncation toward zero, as in the table's builder
  return ((i * i) / 16384) * i / 16384;                                // i <= 26868: both products fit in 31 bits
}

// cv2.cvtColor(COLOR_LAB2RGB) for one uint8 pixel.
TIAB_LAB_HD void lab_to_rgb_u8(const LabTables& t, int L, int a, int b, int& r8, int& g8, int& b8) {
  const int y = t.yf[2 * L], ify = t.yf[2 * L + 1];
  const int adiv = ((5 * a * 53687 + (1 << 7)) >> 13) - 128 * 16384 / 500;       // (a - 128) / 500 in units of 1/16384
  const int bdiv = ((b * 41943 + (1 << 4)) >> 9) - 128 * 16384 / 200 + 1;        // (b - 128) / 200
  const int x = lab_finv_q14(ify + adiv);
  const int z = lab_finv_q14(ify - bdiv);
  // XYZ -> linear sRGB (matrix times white point, scale 4096), result in units of 1/4096
  int ro = (12615 * x - 6296 * y - 2223 * z + (1 << 13)) >> 14;
  int go = (-3773 * x + 7684 * y + 185 * z + (1 << 13)) >> 14;
  int bo = (217 * x - 836 * y + 4715 * z + (1 << 13)) >> 14;
  ro = ro < 0 ? 0 : (ro > 4095 ? 4095 : ro);
  go = go < 0 ? 0 : (go > 4095 ? 4095 : go);
  bo = bo < 0 ? 0 : (bo > 4095 ? 4095 : bo);
  r8 = t.inv_gamma[ro];
  g8 = t.inv_gamma[go];
  b8 = t.inv_gamma[bo];
}

// ---- Reinhard statistics and per-level maps (tools/stainnorm.py:262-284, 286-312, 314-335, 337-367) -------------------------------
// lab_split turns the uint8 Lab image into float32 channels: L / 2.55 (divided in float64, `chan1 /= np.asarray(2.55)`), a - 128, b - 128.
TIAB_LAB_HD float reinhard_chan(int c, int level) { return c == 0 ? (float)((double)level / 2.55) : (float)level - 128.0f; }

// Source statistics from the 256-bin histogram of one uint8 Lab channel: cv2.meanStdDev sums the float32 channel in float64
// (mean = s / n, std = sqrt(max(sq / n - mean^2, 0))).  A channel with a single occupied level has std = 0 exactly (the reference's
// `target_std / std` then raises ZeroDivisionError; cv2's own rounding for constant images is summation-order dependent).
TIAB_LAB_HD void reinhard_stats(int c, const unsigned* hist, double npix, double& mean, double& sd) {
  double s = 0.0, sq = 0.0;
  int levels = 0;
  for (int v = 0; v < 256; ++v) {
    if (!hist[v]) continue;
    const double x = (double)reinhard_chan(c, v), h = (double)hist[v];
    s += h * x;
    sq += h * (x * x);   // x * x is exact in float64 (x is a float32)
    ++levels;
  }
  mean = s / npix;
  const double var = sq / npix - mean * mean;
  sd = (levels > 1 && var > 0.0) ? sqrt(var) : 0.0;
}

// IEEE single-precision operations without FMA contraction (numpy evaluates the reference's expressions one ufunc at a time)
#ifdef __CUDA_ARCH__
TIAB_LAB_HD float lab_fadd(float a, float b) { return __fadd_rn(a, b); }
TIAB_LAB_HD float lab_fsub(float a, float b) { return __fsub_rn(a, b); }
TIAB_LAB_HD float lab_fmul(float a, float b) { return __fmul_rn(a, b); }
#else
TIAB_LAB_HD float lab_fadd(float a, float b) { volatile float r = a + b; return r; }
TIAB_LAB_HD float lab_fsub(float a, float b) { volatile float r = a - b; return r; }
TIAB_LAB_HD float lab_fmul(float a, float b) { volatile float r = a * b; return r; }
#endif

// ReinhardNormalizer.transform + merge_back for one channel level (stainnorm.py:276-284, 329-333): float32 array arithmetic with
// Python-float scalars (cast to float32 first, NEP 50), then clip to [0, 255] and astype(uint8) = truncation.
TIAB_LAB_HD int reinhard_level(int c, int level, double src_mean, double src_sd, double tgt_mean, double tgt_sd) {
  const float x = reinhard_chan(c, level);
  const float n = lab_fadd(lab_fmul(lab_fsub(x, (float)src_mean), (float)(tgt_sd / src_sd)), (float)tgt_mean);
  float m = c == 0 ? lab_fmul(n, 2.55f) : lab_fadd(n, 128.0f);
  m = m < 0.0f ? 0.0f : (m > 255.0f ? 255.0f : m);   // np.clip (NaN cannot occur: src_sd > 0 is checked by the caller)
  return (int)m;
}

}  // namespace tiab
