// reinhard.cu — ReinhardNormalizer on the device (tiatoolbox/tools/stainnorm.py:222-367).
//
// The reference converts the uint8 image to 8-bit Lab with OpenCV, shifts / scales each float32 channel to the target's mean and
// standard deviation, truncates back to uint8 and converts to RGB with OpenCV.  Every step between the two colour conversions is a
// function of the 8-bit channel level alone, so on the device the whole transform is
//     RGB --(fixed-point, lab_math.cuh)--> Lab8 --(three 256-entry level maps per image)--> Lab8' --(fixed-point)--> RGB
// and the per-image statistics come from three 256-bin histograms.  Integer work throughout: HBM-bound, algorithmic bytes per image =
// 2 reads (statistics pass, transform pass) + 1 write of H*W*3 bytes; the tables (10.5 KB) are staged in shared memory because their
// lookups are data dependent.
//
//   reinhard_hist_kernel      grid (chunks, N)  per-block shared histograms of L, a, b -> atomicAdd into hist[N][3][256]
//   reinhard_finalize_kernel  grid N            means / stds (float64, cv2.meanStdDev's formula) and the three level maps of each image
//   reinhard_apply_kernel     grid (chunks, N)  per pixel: RGB -> Lab -> level maps -> RGB
#include <cuda_runtime.h>
#include <math.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/tiab200.h"
#include "ctx.h"
#include "lab_math.cuh"
#include "lab_tables.h"

namespace tiab {

struct __align__(16) LabTablesDev {   // device copies of lab_tables.h (one per process and device, built on first use); read as uint4
  uint16_t gamma[256];
  uint16_t cbrt[2048 + 8];   // 2041 used
  uint16_t yf[512];
  uint8_t inv_gamma[4096];
};

__device__ LabTablesDev g_lab_tables;

struct __align__(16) LabShared {
  uint16_t gamma[256];
  uint16_t cbrt[2048 + 8];
  uint16_t yf[512];
  uint8_t inv_gamma[4096];
};

__device__ __forceinline__ void stage_tables(LabShared& s, bool forward, bool inverse) {
  // 16-byte copies; the struct layouts are identical
  const uint4* src = reinterpret_cast<const uint4*>(&g_lab_tables);
  uint4* dst = reinterpret_cast<uint4*>(&s);
  constexpr int fwd_end = (sizeof(uint16_t) * (256 + 2048 + 8)) / 16;       // gamma + cbrt
  constexpr int total = sizeof(LabTablesDev) / 16;
  static_assert(sizeof(LabTablesDev) % 16 == 0 && sizeof(LabTablesDev) == sizeof(LabShared), "table block is copied in 16-byte words");
  const int lo = forward ? 0 : fwd_end, hi = inverse ? total : fwd_end;
  for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) dst[i] = src[i];
}

constexpr int RH_THREADS = 256;
constexpr int RH_PIX_PER_BLOCK = 8192;

// ---- pass 1: histograms of the 8-bit Lab channels --------------------------------------------------------------------------------
__global__ void __launch_bounds__(RH_THREADS) reinhard_hist_kernel(const uint8_t* __restrict__ in, long long npix,
                                                                   unsigned* __restrict__ hist) {
  __shared__ __align__(16) LabShared tabs;
  __shared__ unsigned h[768];
  stage_tables(tabs, true, false);
  for (int i = threadIdx.x; i < 768; i += RH_THREADS) h[i] = 0u;
  __syncthreads();
  const LabTables t{tabs.gamma, tabs.cbrt, tabs.yf, tabs.inv_gamma};
  const uint8_t* img = in + (long long)blockIdx.y * npix * 3;
  const long long p0 = (long long)blockIdx.x * RH_PIX_PER_BLOCK;
  const long long p1 = min(p0 + RH_PIX_PER_BLOCK, npix);
  const bool vec = ((npix * 3) % 4 == 0) && ((reinterpret_cast<uintptr_t>(in) & 3) == 0);
  if (vec) {
    // 4 pixels = 12 bytes = three aligned 32-bit words per thread and step (p0 is a multiple of 4)
    const long long q1 = p0 + ((p1 - p0) & ~3LL);
    for (long long p = p0 + 4LL * threadIdx.x; p < q1; p += 4LL * RH_THREADS) {
      const uint32_t* w = reinterpret_cast<const uint32_t*>(img + p * 3);
      const uint32_t w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2);
      const int c[12] = {(int)(w0 & 255), (int)((w0 >> 8) & 255), (int)((w0 >> 16) & 255), (int)(w0 >> 24),
                         (int)(w1 & 255), (int)((w1 >> 8) & 255), (int)((w1 >> 16) & 255), (int)(w1 >> 24),
                         (int)(w2 & 255), (int)((w2 >> 8) & 255), (int)((w2 >> 16) & 255), (int)(w2 >> 24)};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        int L, a, b;
        rgb_to_lab_u8(t, c[3 * k], c[3 * k + 1], c[3 * k + 2], L, a, b);
        atomicAdd(&h[L], 1u);
        atomicAdd(&h[256 + a], 1u);
        atomicAdd(&h[512 + b], 1u);
      }
    }
    for (long long p = q1 + threadIdx.x; p < p1; p += RH_THREADS) {
      int L, a, b;
      rgb_to_lab_u8(t, img[p * 3], img[p * 3 + 1], img[p * 3 + 2], L, a, b);
      atomicAdd(&h[L], 1u);
      atomicAdd(&h[256 + a], 1u);
      atomicAdd(&h[512 + b], 1u);
    }
  } else {
    for (long long p = p0 + threadIdx.x; p < p1; p += RH_THREADS) {
      int L, a, b;
      rgb_to_lab_u8(t, img[p * 3], img[p * 3 + 1], img[p * 3 + 2], L, a, b);
      atomicAdd(&h[L], 1u);
      atomicAdd(&h[256 + a], 1u);
      atomicAdd(&h[512 + b], 1u);
    }
  }
  __syncthreads();
  unsigned* g = hist + (long long)blockIdx.y * 768;
  for (int i = threadIdx.x; i < 768; i += RH_THREADS)
    if (h[i]) atomicAdd(&g[i], h[i]);
}

// ---- pass 2: statistics + level maps ---------------------------------------------------------------------------------------------
// stats[n] = {mean_L, mean_a, mean_b, sd_L, sd_a, sd_b}; status[n] = 1 when a channel is constant (the reference's
// `target_std / std` then raises ZeroDivisionError); lut[n][3][256] (identity for status != 0 so that the image passes through).
__global__ void __launch_bounds__(256) reinhard_finalize_kernel(const unsigned* __restrict__ hist, double npix, const double* tgt,
                                                                double* __restrict__ stats, int* __restrict__ status,
                                                                uint8_t* __restrict__ lut) {
  __shared__ double s_stats[6];
  const unsigned* h = hist + (long long)blockIdx.x * 768;
  if (threadIdx.x < 3) {
    double m, sd;
    reinhard_stats(threadIdx.x, h + 256 * threadIdx.x, npix, m, sd);
    s_stats[threadIdx.x] = m;
    s_stats[3 + threadIdx.x] = sd;
  }
  __syncthreads();
  const bool bad = !(s_stats[3] > 0.0) || !(s_stats[4] > 0.0) || !(s_stats[5] > 0.0);
  if (threadIdx.x < 6) stats[(long long)blockIdx.x * 6 + threadIdx.x] = s_stats[threadIdx.x];
  if (threadIdx.x == 0) status[blockIdx.x] = bad ? 1 : 0;
  if (lut == nullptr) return;
  uint8_t* out = lut + (long long)blockIdx.x * 768;
  const int v = threadIdx.x;
  for (int c = 0; c < 3; ++c)
    out[256 * c + v] = bad ? (uint8_t)v : (uint8_t)reinhard_level(c, v, s_stats[c], s_stats[3 + c], tgt[c], tgt[3 + c]);
}

// ---- pass 3: per-pixel transform ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void reinhard_pixel(const LabTables& t, const uint8_t* m, int r, int g, int b, int& ro, int& go, int& bo) {
  int L, a, bb;
  rgb_to_lab_u8(t, r, g, b, L, a, bb);
  lab_to_rgb_u8(t, m[L], m[256 + a], m[512 + bb], ro, go, bo);
}

__global__ void __launch_bounds__(RH_THREADS) reinhard_apply_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out,
                                                                    long long npix, const uint8_t* __restrict__ lut,
                                                                    const int* __restrict__ status) {
  __shared__ __align__(16) LabShared tabs;
  __shared__ __align__(16) uint8_t m[768];
  const long long img_off = (long long)blockIdx.y * npix * 3;
  const long long p0 = (long long)blockIdx.x * RH_PIX_PER_BLOCK;
  const long long p1 = min(p0 + RH_PIX_PER_BLOCK, npix);
  const uint8_t* img = in + img_off;
  uint8_t* dst = out + img_off;
  if (status[blockIdx.y] != 0) {   // constant channel: pass the image through (the host side raises like the reference)
    for (long long i = p0 * 3 + threadIdx.x; i < p1 * 3; i += RH_THREADS) dst[i] = img[i];
    return;
  }
  stage_tables(tabs, true, true);
  for (int i = threadIdx.x; i < 768 / 4; i += RH_THREADS)
    reinterpret_cast<uint32_t*>(m)[i] = reinterpret_cast<const uint32_t*>(lut + (long long)blockIdx.y * 768)[i];
  __syncthreads();
  const LabTables t{tabs.gamma, tabs.cbrt, tabs.yf, tabs.inv_gamma};
  const bool vec = ((npix * 3) % 4 == 0) && ((reinterpret_cast<uintptr_t>(in) & 3) == 0) && ((reinterpret_cast<uintptr_t>(out) & 3) == 0);
  long long q1 = p0;
  if (vec) {
    q1 = p0 + ((p1 - p0) & ~3LL);
    for (long long p = p0 + 4LL * threadIdx.x; p < q1; p += 4LL * RH_THREADS) {
      const uint32_t* w = reinterpret_cast<const uint32_t*>(img + p * 3);
      const uint32_t w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2);
      const int c[12] = {(int)(w0 & 255), (int)((w0 >> 8) & 255), (int)((w0 >> 16) & 255), (int)(w0 >> 24),
                         (int)(w1 & 255), (int)((w1 >> 8) & 255), (int)((w1 >> 16) & 255), (int)(w1 >> 24),
                         (int)(w2 & 255), (int)((w2 >> 8) & 255), (int)((w2 >> 16) & 255), (int)(w2 >> 24)};
      int o[12];
#pragma unroll
      for (int k = 0; k < 4; ++k) reinhard_pixel(t, m, c[3 * k], c[3 * k + 1], c[3 * k + 2], o[3 * k], o[3 * k + 1], o[3 * k + 2]);
      uint32_t* wo = reinterpret_cast<uint32_t*>(dst + p * 3);
      wo[0] = (uint32_t)o[0] | ((uint32_t)o[1] << 8) | ((uint32_t)o[2] << 16) | ((uint32_t)o[3] << 24);
      wo[1] = (uint32_t)o[4] | ((uint32_t)o[5] << 8) | ((uint32_t)o[6] << 16) | ((uint32_t)o[7] << 24);
      wo[2] = (uint32_t)o[8] | ((uint32_t)o[9] << 8) | ((uint32_t)o[10] << 16) | ((uint32_t)o[11] << 24);
    }
  }
  for (long long p = q1 + threadIdx.x; p < p1; p += RH_THREADS) {
    int ro, go, bo;
    reinhard_pixel(t, m, img[p * 3], img[p * 3 + 1], img[p * 3 + 2], ro, go, bo);
    dst[p * 3] = (uint8_t)ro;
    dst[p * 3 + 1] = (uint8_t)go;
    dst[p * 3 + 2] = (uint8_t)bo;
  }
}

// ---- uint8 colour conversion on its own (cv2.cvtColor on the device; also the unit-test surface of lab_math.cuh) -------------------
__global__ void __launch_bounds__(RH_THREADS) lab_convert_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, long long npix,
                                                                 int inverse) {
  __shared__ __align__(16) LabShared tabs;
  stage_tables(tabs, inverse == 0, inverse != 0);
  __syncthreads();
  const LabTables t{tabs.gamma, tabs.cbrt, tabs.yf, tabs.inv_gamma};
  for (long long p = (long long)blockIdx.x * RH_THREADS + threadIdx.x; p < npix; p += (long long)gridDim.x * RH_THREADS) {
    int x, y, z;
    if (inverse) lab_to_rgb_u8(t, in[p * 3], in[p * 3 + 1], in[p * 3 + 2], x, y, z);
    else rgb_to_lab_u8(t, in[p * 3], in[p * 3 + 1], in[p * 3 + 2], x, y, z);
    out[p * 3] = (uint8_t)x;
    out[p * 3 + 1] = (uint8_t)y;
    out[p * 3 + 2] = (uint8_t)z;
  }
}

// ---- experimental: the three passes fused, one CTA per image (opt-in: tiab_reinhard_set_path(h, 1)) ---------------------------------------
// For images of at most 192 KB (256 x 256 x 3): the image is read ONCE into shared memory, converted to Lab8 in place while the three
// histograms are built, the level maps are formed by the same finalise code, and the second pass reads Lab from shared memory — one
// forward conversion instead of two and 2 x H x W x 3 bytes of HBM traffic instead of 3 x.  Persistent CTAs (one per SM) loop over images.
constexpr int RF_THREADS = 1024;
constexpr int RF_MAX_BYTES = 196608;

__global__ void __launch_bounds__(RF_THREADS) reinhard_fused_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int N,
                                                                    int npix, const double* __restrict__ tgt, double* __restrict__ stats,
                                                                    int* __restrict__ status) {
  extern __shared__ __align__(16) unsigned char rf_smem[];
  LabShared& tabs = *reinterpret_cast<LabShared*>(rf_smem);
  unsigned* h = reinterpret_cast<unsigned*>(rf_smem + sizeof(LabShared));            // [768]
  uint8_t* m = reinterpret_cast<uint8_t*>(h + 768);                                  // [768]
  double* s_stats = reinterpret_cast<double*>(m + 768);                              // [6] (+ flag)
  uint8_t* img = reinterpret_cast<uint8_t*>(s_stats + 8);                            // [npix * 3], 16-byte aligned
  const int nbytes = npix * 3, nvec = nbytes >> 4, nquad = npix >> 2;                // the caller guarantees nbytes % 16 == 0
  stage_tables(tabs, true, true);
  const LabTables t{tabs.gamma, tabs.cbrt, tabs.yf, tabs.inv_gamma};
  for (int n = blockIdx.x; n < N; n += gridDim.x) {
    const uint4* src = reinterpret_cast<const uint4*>(in + (long long)n * nbytes);
    uint4* dst = reinterpret_cast<uint4*>(out + (long long)n * nbytes);
    for (int i = threadIdx.x; i < nvec; i += RF_THREADS) reinterpret_cast<uint4*>(img)[i] = __ldg(src + i);
    for (int i = threadIdx.x; i < 768; i += RF_THREADS) h[i] = 0u;
    __syncthreads();
    // pass 1: RGB -> Lab8 in place (4 pixels = three 32-bit words per thread and step) + histograms
    for (int q = threadIdx.x; q < nquad; q += RF_THREADS) {
      uint32_t* w = reinterpret_cast<uint32_t*>(img) + 3 * q;
      const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
      const int c[12] = {(int)(w0 & 255), (int)((w0 >> 8) & 255), (int)((w0 >> 16) & 255), (int)(w0 >> 24),
                         (int)(w1 & 255), (int)((w1 >> 8) & 255), (int)((w1 >> 16) & 255), (int)(w1 >> 24),
                         (int)(w2 & 255), (int)((w2 >> 8) & 255), (int)((w2 >> 16) & 255), (int)(w2 >> 24)};
      int o[12];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        rgb_to_lab_u8(t, c[3 * k], c[3 * k + 1], c[3 * k + 2], o[3 * k], o[3 * k + 1], o[3 * k + 2]);
        atomicAdd(&h[o[3 * k]], 1u);
        atomicAdd(&h[256 + o[3 * k + 1]], 1u);
        atomicAdd(&h[512 + o[3 * k + 2]], 1u);
      }
      w[0] = (uint32_t)o[0] | ((uint32_t)o[1] << 8) | ((uint32_t)o[2] << 16) | ((uint32_t)o[3] << 24);
      w[1] = (uint32_t)o[4] | ((uint32_t)o[5] << 8) | ((uint32_t)o[6] << 16) | ((uint32_t)o[7] << 24);
      w[2] = (uint32_t)o[8] | ((uint32_t)o[9] << 8) | ((uint32_t)o[10] << 16) | ((uint32_t)o[11] << 24);
    }
    __syncthreads();
    // statistics + level maps (the code of reinhard_finalize_kernel)
    if (threadIdx.x < 3) {
      double mean, sd;
      reinhard_stats(threadIdx.x, h + 256 * threadIdx.x, (double)npix, mean, sd);
      s_stats[threadIdx.x] = mean;
      s_stats[3 + threadIdx.x] = sd;
    }
    __syncthreads();
    const bool bad = !(s_stats[3] > 0.0) || !(s_stats[4] > 0.0) || !(s_stats[5] > 0.0);
    if (threadIdx.x < 6) stats[(long long)n * 6 + threadIdx.x] = s_stats[threadIdx.x];
    if (threadIdx.x == 0) status[n] = bad ? 1 : 0;
    if (bad) {   // constant channel: the image passes through unchanged (block-uniform branch)
      for (int i = threadIdx.x; i < nvec; i += RF_THREADS) dst[i] = __ldg(src + i);
      __syncthreads();
      continue;
    }
    if (threadIdx.x < 256)
      for (int c = 0; c < 3; ++c)
        m[256 * c + threadIdx.x] = (uint8_t)reinhard_level(c, threadIdx.x, s_stats[c], s_stats[3 + c], tgt[c], tgt[3 + c]);
    __syncthreads();
    // pass 2: Lab8 -> level maps -> RGB in place, then coalesced 16-byte stores
    for (int q = threadIdx.x; q < nquad; q += RF_THREADS) {
      uint32_t* w = reinterpret_cast<uint32_t*>(img) + 3 * q;
      const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
      const int c[12] = {(int)(w0 & 255), (int)((w0 >> 8) & 255), (int)((w0 >> 16) & 255), (int)(w0 >> 24),
                         (int)(w1 & 255), (int)((w1 >> 8) & 255), (int)((w1 >> 16) & 255), (int)(w1 >> 24),
                         (int)(w2 & 255), (int)((w2 >> 8) & 255), (int)((w2 >> 16) & 255), (int)(w2 >> 24)};
      int o[12];
#pragma unroll
      for (int k = 0; k < 4; ++k)
        lab_to_rgb_u8(t, m[c[3 * k]], m[256 + c[3 * k + 1]], m[512 + c[3 * k + 2]], o[3 * k], o[3 * k + 1], o[3 * k + 2]);
      w[0] = (uint32_t)o[0] | ((uint32_t)o[1] << 8) | ((uint32_t)o[2] << 16) | ((uint32_t)o[3] << 24);
      w[1] = (uint32_t)o[4] | ((uint32_t)o[5] << 8) | ((uint32_t)o[6] << 16) | ((uint32_t)o[7] << 24);
      w[2] = (uint32_t)o[8] | ((uint32_t)o[9] << 8) | ((uint32_t)o[10] << 16) | ((uint32_t)o[11] << 24);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nvec; i += RF_THREADS) dst[i] = reinterpret_cast<const uint4*>(img)[i];
    __syncthreads();
  }
}

static int ensure_tables(tiab_ctx* h, cudaStream_t s) {
  if (h->lab_tables_set) return TIAB_OK;
  static LabTablesDev host;   // zero-initialised padding
  std::copy(kSrgbGamma2040, kSrgbGamma2040 + 256, host.gamma);
  std::copy(kLabCbrt, kLabCbrt + 2041, host.cbrt);
  std::copy(kLabToYF, kLabToYF + 512, host.yf);
  std::copy(kSrgbInvGamma, kSrgbInvGamma + 4096, host.inv_gamma);
  TIAB_CUDA_OK(h, cudaMemcpyToSymbolAsync(g_lab_tables, &host, sizeof(host), 0, cudaMemcpyHostToDevice, s));
  TIAB_CUDA_OK(h, cudaStreamSynchronize(s));
  h->lab_tables_set = 1;
  return TIAB_OK;
}

static int launch_ok(tiab_ctx* h, const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    h->err = std::string(what) + ": " + cudaGetErrorString(e);
    return TIAB_E_CUDA;
  }
  return TIAB_OK;
}

// statistics (+ level maps when tgt_dev != nullptr) of N images; everything stays on the device
static int reinhard_stats_dev(tiab_ctx* h, const uint8_t* in, int N, long long npix, const double* tgt_dev, unsigned* hist, double* stats,
                              int* status, uint8_t* lut, cudaStream_t s) {
  TIAB_CUDA_OK(h, cudaMemsetAsync(hist, 0, (size_t)N * 768 * sizeof(unsigned), s));
  const int chunks = (int)((npix + RH_PIX_PER_BLOCK - 1) / RH_PIX_PER_BLOCK);
  reinhard_hist_kernel<<<dim3(chunks, N), RH_THREADS, 0, s>>>(in, npix, hist);
  h->last_launches++;
  if (int rc = launch_ok(h, "reinhard_hist_kernel")) return rc;
  reinhard_finalize_kernel<<<N, 256, 0, s>>>(hist, (double)npix, tgt_dev, stats, status, tgt_dev ? lut : nullptr);
  h->last_launches++;
  return launch_ok(h, "reinhard_finalize_kernel");
}

}  // namespace tiab

using namespace tiab;

extern "C" {

int tiab_reinhard_fit(tiab_handle_t h, const uint8_t* target_dev, int H, int W, double means_host[3], double stds_host[3], void* stream) {
  if (!h) return TIAB_E_INVALID;
  if (!target_dev || H <= 0 || W <= 0 || !means_host || !stds_host) {
    h->err = "tiab_reinhard_fit: invalid argument";
    return TIAB_E_INVALID;
  }
  cudaStream_t s = (cudaStream_t)stream;
  TIAB_ON_DEVICE(h);
  if (int rc = ensure_tables(h, s)) return rc;
  h->last_launches = 0;
  void* work = nullptr;
  const size_t bytes = 768 * sizeof(unsigned) + 6 * sizeof(double) + 16;
  TIAB_CUDA_OK(h, cudaMallocAsync(&work, bytes, s));
  double* stats = (double*)work;
  unsigned* hist = (unsigned*)(stats + 6);
  int* status = (int*)(hist + 768);
  int rc = reinhard_stats_dev(h, target_dev, 1, (long long)H * W, nullptr, hist, stats, status, nullptr, s);
  double host[6];
  if (rc == TIAB_OK) {
    TIAB_CUDA_OK(h, cudaMemcpyAsync(host, stats, sizeof(host), cudaMemcpyDeviceToHost, s));
    TIAB_CUDA_OK(h, cudaStreamSynchronize(s));
    for (int c = 0; c < 3; ++c) means_host[c] = host[c], stds_host[c] = host[3 + c];
  }
  cudaFreeAsync(work, s);
  return rc;
}

int tiab_reinhard_transform_batch(tiab_handle_t h, const uint8_t* in_dev, uint8_t* out_dev, int N, int H, int W, const double target_means[3],
                                  const double target_stds[3], double* src_stats_dev, int* status_dev, void* stream) {
  if (!h) return TIAB_E_INVALID;
  if (!in_dev || !out_dev || N <= 0 || N > 65535 || H <= 0 || W <= 0 || !target_means || !target_stds || in_dev == out_dev) {
    h->err = "tiab_reinhard_transform_batch: invalid argument (1 <= N <= 65535, distinct in / out buffers)";
    return TIAB_E_INVALID;
  }
  cudaStream_t s = (cudaStream_t)stream;
  TIAB_ON_DEVICE(h);
  if (int rc = ensure_tables(h, s)) return rc;
  h->last_launches = 0;
  const long long npix = (long long)H * W;
  // workspace: tgt[6] doubles | stats[N*6] doubles | hist[N*768] | status[N] | lut[N*768]
  const size_t bytes = (6 + (size_t)N * 6) * sizeof(double) + (size_t)N * 768 * sizeof(unsigned) + (size_t)N * sizeof(int) + (size_t)N * 768 + 64;
  void* work = nullptr;
  TIAB_CUDA_OK(h, cudaMallocAsync(&work, bytes, s));
  double* tgt = (double*)work;
  double* stats = tgt + 6;
  unsigned* hist = (unsigned*)(stats + (size_t)N * 6);
  int* status = (int*)(hist + (size_t)N * 768);
  uint8_t* lut = (uint8_t*)(status + N);
  lut = (uint8_t*)(((uintptr_t)lut + 15) & ~(uintptr_t)15);
  double tgt_host[6] = {target_means[0], target_means[1], target_means[2], target_stds[0], target_stds[1], target_stds[2]};
  TIAB_CUDA_OK(h, cudaMemcpyAsync(tgt, tgt_host, sizeof(tgt_host), cudaMemcpyHostToDevice, s));
  TIAB_CUDA_OK(h, cudaStreamSynchronize(s));   // tgt_host lives on this stack frame
  const bool fused_ok = h->reinhard_path == 1 && npix * 3 <= RF_MAX_BYTES && (npix * 3) % 16 == 0 && npix % 4 == 0 &&
                        ((reinterpret_cast<uintptr_t>(in_dev) | reinterpret_cast<uintptr_t>(out_dev)) & 15) == 0;
  int rc = TIAB_OK;
  if (fused_ok) {
    const size_t smem = sizeof(LabShared) + 768 * sizeof(unsigned) + 768 + 8 * sizeof(double) + (size_t)npix * 3;
    TIAB_CUDA_OK(h, cudaFuncSetAttribute(reinhard_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    reinhard_fused_kernel<<<std::min(N, h->sm_count), RF_THREADS, smem, s>>>(in_dev, out_dev, N, (int)npix, tgt, stats, status);
    h->last_launches++;
    rc = launch_ok(h, "reinhard_fused_kernel");
  } else {
    rc = reinhard_stats_dev(h, in_dev, N, npix, tgt, hist, stats, status, lut, s);
  }
  if (rc == TIAB_OK && !fused_ok) {
    const int chunks = (int)((npix + RH_PIX_PER_BLOCK - 1) / RH_PIX_PER_BLOCK);
    reinhard_apply_kernel<<<dim3(chunks, N), RH_THREADS, 0, s>>>(in_dev, out_dev, npix, lut, status);
    h->last_launches++;
    rc = launch_ok(h, "reinhard_apply_kernel");
  }
  if (rc == TIAB_OK && src_stats_dev)
    TIAB_CUDA_OK(h, cudaMemcpyAsync(src_stats_dev, stats, (size_t)N * 6 * sizeof(double), cudaMemcpyDeviceToDevice, s));
  if (rc == TIAB_OK && status_dev) TIAB_CUDA_OK(h, cudaMemcpyAsync(status_dev, status, (size_t)N * sizeof(int), cudaMemcpyDeviceToDevice, s));
  cudaFreeAsync(work, s);
  return rc;
}

int tiab_reinhard_set_path(tiab_handle_t h, int path) {
  if (!h || (path != 0 && path != 1)) return TIAB_E_INVALID;
  h->reinhard_path = path;
  return TIAB_OK;
}

int tiab_lab_convert(tiab_handle_t h, const uint8_t* in_dev, uint8_t* out_dev, long long npix, int inverse, void* stream) {
  if (!h) return TIAB_E_INVALID;
  if (!in_dev || !out_dev || npix <= 0) {
    h->err = "tiab_lab_convert: invalid argument";
    return TIAB_E_INVALID;
  }
  cudaStream_t s = (cudaStream_t)stream;
  TIAB_ON_DEVICE(h);
  if (int rc = ensure_tables(h, s)) return rc;
  const int grid = (int)std::min<long long>((npix + RH_THREADS - 1) / RH_THREADS, (long long)h->sm_count * 8);
  lab_convert_kernel<<<grid, RH_THREADS, 0, s>>>(in_dev, out_dev, npix, inverse);
  h->last_launches = 1;
  return launch_ok(h, "lab_convert_kernel");
}

// ---- host-side hooks for the CPU unit tests (no GPU): the same lab_math.cuh functions compiled for the host -------------------------
int tiab_dbg_lab_convert(const uint8_t* in_host, uint8_t* out_host, long long npix, int inverse) {
  if (!in_host || !out_host || npix < 0) return TIAB_E_INVALID;
  const LabTables t{kSrgbGamma2040, kLabCbrt, kLabToYF, kSrgbInvGamma};
  for (long long p = 0; p < npix; ++p) {
    int x, y, z;
    if (inverse) lab_to_rgb_u8(t, in_host[p * 3], in_host[p * 3 + 1], in_host[p * 3 + 2], x, y, z);
    else rgb_to_lab_u8(t, in_host[p * 3], in_host[p * 3 + 1], in_host[p * 3 + 2], x, y, z);
    out_host[p * 3] = (uint8_t)x;
    out_host[p * 3 + 1] = (uint8_t)y;
    out_host[p * 3 + 2] = (uint8_t)z;
  }
  return TIAB_OK;
}

int tiab_dbg_reinhard_maps(const unsigned hist_host[768], double npix, const double target_means[3], const double target_stds[3],
                           double stats_out[6], uint8_t lut_out[768]) {
  if (!hist_host || !target_means || !target_stds || !stats_out || !lut_out || !(npix > 0)) return TIAB_E_INVALID;
  for (int c = 0; c < 3; ++c) reinhard_stats(c, hist_host + 256 * c, npix, stats_out[c], stats_out[3 + c]);
  const bool bad = !(stats_out[3] > 0.0) || !(stats_out[4] > 0.0) || !(stats_out[5] > 0.0);
  for (int c = 0; c < 3; ++c)
    for (int v = 0; v < 256; ++v)
      lut_out[256 * c + v] = bad ? (uint8_t)v : (uint8_t)reinhard_level(c, v, stats_out[c], stats_out[3 + c], target_means[c], target_stds[c]);
  return bad ? 1 : TIAB_OK;
}

}  // extern "C"
