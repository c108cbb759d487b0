// postproc.cu — HoVer-Net post-processing on the device (sm_100a).
//
// Reference: HoVerNet._proc_np_hv + get_instance_info, tiatoolbox/models/architecture/hovernet.py:502-748
// (CPU: scipy.ndimage.label / binary_fill_holes, cv2.normalize / Sobel(ksize=21, CV_64F) / GaussianBlur /
// morphologyEx, skimage remove_small_objects / watershed).  Every stage is restated as a kernel over
// all pixels of all maps of the batch:
//
//   threshold -> CCL (union-find, ids in raster order of the first pixel = ndimage.label) -> drop <= 9 px
//   min-max normalise h, v -> separable 21-tap Sobel in float64 (REFLECT_101, OpenCV's summation order)
//   -> min-max normalise -> energy / distance maps -> 3x3 Gaussian (float64) -> marker seeds
//   -> fill holes (background CCL not touching the border) -> 5x5 elliptic opening -> CCL -> drop <= 9 px
//   -> marker-controlled watershed: one thread per foreground component runs the sequential priority
//      flood (key = (value, insertion age), label on push, neighbours -row,-1,+1,+row) in a private heap
//   -> per-instance table (box, centroid, majority type, type probability).
//
// This is integer / index work bounded by latency, not bandwidth (20 B/px algorithmic).
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/tiab200.h"
#include "contour_math.cuh"
#include "ctx.h"
#include "watershed_math.cuh"

namespace tiab {

constexpr int PP_THREADS = 256;
constexpr int PP_MAX_TYPES = 16;

struct HeapItem {
  double v;
  unsigned age;
  unsigned idx;
};

struct PPWork {  // all arrays are [N * P] unless noted
  int* lab;            // union-find labels / component root (local pixel index), -1 background
  int* size;           // pixels per root
  int* cid;            // raster-order component id (1-based) stored at the root pixel
  int* roots;          // [N * P] list of roots in raster order
  int* ncomp;          // [N]
  int* bbox;           // [N * P * 4] per-root bounding box (x0, y0, x1, y1 inclusive), used for foreground comps
  uint8_t* blb;        // foreground after small-object removal
  uint8_t* fg;         // scratch binary map (CCL input)
  uint8_t* mk;         // marker seeds (binary)
  uint8_t* tmp;        // morphology scratch
  uint8_t* border;     // [N * P] per-root "touches the border" flag (fill holes)
  float* hn;           // normalised h / later 1 - normalised sobel_h
  float* vn;
  double* rowh;        // row-filtered
  double* rowv;
  double* sobh;
  double* sobv;
  double* dist0;       // (1 - overall) * blb
  double* dist;        // -GaussianBlur(dist0)
  float* mm;           // [N * 8] min/max scratch: h min,max, v min,max ; as doubles in mm64
  double* mm64;        // [N * 4] sobel_h min,max, sobel_v min,max
  int* marker;         // labelled markers (ids kept, small removed)
  HeapItem* heap;      // [N * P]
  int* heap_off;       // [N * P] heap slice offset per foreground root
  // instance accumulators, [N * maxI ...]
  int* iarea;
  int* ibox;           // [N*maxI*4]
  unsigned long long* isx;
  unsigned long long* isy;
  int* itype;          // [N*maxI*PP_MAX_TYPES]
};

__constant__ double c_sobel_d[21];     // ksize 21 (scale_factor 1)
__constant__ double c_sobel_s[21];
__constant__ double c_sobel_d11[11];   // ksize 11 (scale_factor 0.5: HoVerNet+, hovernetplus.py:305)
__constant__ double c_sobel_s11[11];
template <int K> __device__ __forceinline__ const double* sobel_d() { return K == 21 ? c_sobel_d : c_sobel_d11; }
template <int K> __device__ __forceinline__ const double* sobel_s() { return K == 21 ? c_sobel_s : c_sobel_s11; }

__device__ __forceinline__ int refl101(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * n - 2 - i;
  return i;
}

// ---- union-find ----
__device__ __forceinline__ int uf_find(const int* L, int i) {
  int p = L[i];
  while (p != i) {
    i = p;
    p = L[i];
  }
  return i;
}
__device__ __forceinline__ void uf_union(int* L, int a, int b) {
  bool done;
  do {
    a = uf_find(L, a);
    b = uf_find(L, b);
    if (a < b) {
      int old = atomicMin(&L[b], a);
      done = (old == b);
      b = old;
    } else if (b < a) {
      int old = atomicMin(&L[a], b);
      done = (old == a);
      a = old;
    } else {
      done = true;
    }
  } while (!done);
}

#define PP_GRID_LOOP(i, total) \
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (total); i += (long long)gridDim.x * blockDim.x)

__global__ void pp_threshold(const float* __restrict__ npm, long long total, uint8_t* __restrict__ fg) {
  PP_GRID_LOOP(i, total) fg[i] = npm[i] >= 0.5f ? 1 : 0;
}
__global__ void pp_ccl_init(const uint8_t* __restrict__ fg, int P, int W, long long total, int* __restrict__ lab, int* __restrict__ size,
                            int* __restrict__ bbox, uint8_t* __restrict__ border) {
  PP_GRID_LOOP(i, total) {
    const int p = (int)(i % P);
    // initial label = first pixel of the pixel's horizontal run (the run is one set from the start: no horizontal unions needed)
    int s = -1;
    if (fg[i]) {
        const int x = p % W;
        int k = 0;
        while (k < x && fg[i - k - 1]) ++k;
        s = p - k;
    }
    lab[i] = s;
    size[i] = 0;
    border[i] = 0;
    if (bbox) {
      bbox[4 * i] = 0x7fffffff, bbox[4 * i + 1] = 0x7fffffff, bbox[4 * i + 2] = -1, bbox[4 * i + 3] = -1;
    }
  }
}
__global__ void pp_ccl_merge(const uint8_t* __restrict__ fg, int H, int W, long long total, int* __restrict__ lab) {
  const int P = H * W;
  PP_GRID_LOOP(i, total) {
    if (!fg[i]) continue;
    const int p = (int)(i % P);
    const long long base = i - p;
    const int x = p % W, y = p / W;
    int* L = lab + base;
    // two vertically adjacent runs are joined once, at the first column they share (the start of the upper or of the lower run)
    if (y + 1 < H && fg[i + W] && (x == 0 || !fg[i - 1] || !fg[i + W - 1])) uf_union(L, p, p + W);
  }
}
__global__ void pp_ccl_flatten(int H, int W, long long total, int* __restrict__ lab, int* __restrict__ size, int* __restrict__ bbox,
                               uint8_t* __restrict__ border) {
  const int P = H * W;
  PP_GRID_LOOP(i, total) {
    const int p = (int)(i % P);
    if (lab[i] < 0) continue;
    const long long base = i - p;
    const int r = uf_find(lab + base, p);
    lab[i] = r;  // safe: a root keeps L[r] == r, non-roots only ever point to smaller indices
    atomicAdd(&size[base + r], 1);
    const int x = p % W, y = p / W;
    if (bbox) {
      int* b = bbox + 4 * (base + r);
      atomicMin(&b[0], x), atomicMin(&b[1], y), atomicMax(&b[2], x), atomicMax(&b[3], y);
    }
    if (x == 0 || y == 0 || x == W - 1 || y == H - 1) border[base + r] = 1;
  }
}
// raster-order ids of the roots: one CTA per image
__global__ void __launch_bounds__(1024) pp_rank_roots(const int* __restrict__ lab, int P, int* __restrict__ cid, int* __restrict__ roots,
                                                      int* __restrict__ ncomp) {
  __shared__ int s_cnt[1024];
  const int n = blockIdx.x;
  const long long base = (long long)n * P;
  const int chunk = (P + 1023) / 1024;
  const int p0 = threadIdx.x * chunk, p1 = min(P, p0 + chunk);
  int c = 0;
  for (int p = p0; p < p1; ++p) c += (lab[base + p] == p) ? 1 : 0;
  s_cnt[threadIdx.x] = c;
  __syncthreads();
  // inclusive scan (Hillis-Steele)
  for (int d = 1; d < 1024; d <<= 1) {
    int v = threadIdx.x >= d ? s_cnt[threadIdx.x - d] : 0;
    __syncthreads();
    s_cnt[threadIdx.x] += v;
    __syncthreads();
  }
  int run = s_cnt[threadIdx.x] - c;
  for (int p = p0; p < p1; ++p)
    if (lab[base + p] == p) {
      roots[base + run] = p;
      cid[base + p] = ++run;
    }
  if (threadIdx.x == 1023) ncomp[n] = s_cnt[1023];
}
// keep[p] = component has more than max_size pixels ; optionally write labelled output (id or 0)
__global__ void pp_keep_large(const int* __restrict__ lab, const int* __restrict__ size, const int* __restrict__ cid, int P, long long total,
                              int max_size, uint8_t* __restrict__ keep, int* __restrict__ labelled) {
  PP_GRID_LOOP(i, total) {
    const int r = lab[i];
    const long long base = i - (i % P);
    const bool k = r >= 0 && size[base + r] > max_size;
    if (keep) keep[i] = k ? 1 : 0;
    if (labelled) labelled[i] = k ? cid[base + r] : 0;
  }
}

// ---- min / max per image (float and double variants) ----
__device__ __forceinline__ void atomic_min_f(float* a, float v) {
  // works for any sign: compare-and-swap loop
  float old = *a;
  while (v < old) {
    int assumed = __float_as_int(old);
    int prev = atomicCAS(reinterpret_cast<int*>(a), assumed, __float_as_int(v));
    if (prev == assumed) break;
    old = __int_as_float(prev);
  }
}
__device__ __forceinline__ void atomic_max_f(float* a, float v) {
  float old = *a;
  while (v > old) {
    int assumed = __float_as_int(old);
    int prev = atomicCAS(reinterpret_cast<int*>(a), assumed, __float_as_int(v));
    if (prev == assumed) break;
    old = __int_as_float(prev);
  }
}
__device__ __forceinline__ void atomic_min_d(double* a, double v) {
  double old = *a;
  while (v < old) {
    unsigned long long assumed = __double_as_longlong(old);
    unsigned long long prev = atomicCAS(reinterpret_cast<unsigned long long*>(a), assumed, (unsigned long long)__double_as_longlong(v));
    if (prev == assumed) break;
    old = __longlong_as_double(prev);
  }
}
__device__ __forceinline__ void atomic_max_d(double* a, double v) {
  double old = *a;
  while (v > old) {
    unsigned long long assumed = __double_as_longlong(old);
    unsigned long long prev = atomicCAS(reinterpret_cast<unsigned long long*>(a), assumed, (unsigned long long)__double_as_longlong(v));
    if (prev == assumed) break;
    old = __longlong_as_double(prev);
  }
}
__global__ void pp_mm_init(int N, float* mm, double* mm64) {
  PP_GRID_LOOP(i, N) {
    mm[4 * i] = FLT_MAX, mm[4 * i + 1] = -FLT_MAX, mm[4 * i + 2] = FLT_MAX, mm[4 * i + 3] = -FLT_MAX;
    mm64[4 * i] = DBL_MAX, mm64[4 * i + 1] = -DBL_MAX, mm64[4 * i + 2] = DBL_MAX, mm64[4 * i + 3] = -DBL_MAX;
  }
}
// grid = (blocks, N)
__global__ void __launch_bounds__(PP_THREADS) pp_minmax_hv(const float* __restrict__ hv, int P, float* __restrict__ mm) {
  const int n = blockIdx.y;
  float a = FLT_MAX, b = -FLT_MAX, c = FLT_MAX, d = -FLT_MAX;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < P; p += gridDim.x * blockDim.x) {
    const float h = hv[((long long)n * P + p) * 2], v = hv[((long long)n * P + p) * 2 + 1];
    a = fminf(a, h), b = fmaxf(b, h), c = fminf(c, v), d = fmaxf(d, v);
  }
  for (int o = 16; o; o >>= 1) {
    a = fminf(a, __shfl_xor_sync(0xffffffffu, a, o)), b = fmaxf(b, __shfl_xor_sync(0xffffffffu, b, o));
    c = fminf(c, __shfl_xor_sync(0xffffffffu, c, o)), d = fmaxf(d, __shfl_xor_sync(0xffffffffu, d, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomic_min_f(&mm[4 * n], a), atomic_max_f(&mm[4 * n + 1], b), atomic_min_f(&mm[4 * n + 2], c), atomic_max_f(&mm[4 * n + 3], d);
  }
}
__global__ void __launch_bounds__(PP_THREADS) pp_minmax_sobel(const double* __restrict__ sh, const double* __restrict__ sv, int P,
                                                              double* __restrict__ mm64) {
  const int n = blockIdx.y;
  double a = DBL_MAX, b = -DBL_MAX, c = DBL_MAX, d = -DBL_MAX;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < P; p += gridDim.x * blockDim.x) {
    const double h = sh[(long long)n * P + p], v = sv[(long long)n * P + p];
    a = fmin(a, h), b = fmax(b, h), c = fmin(c, v), d = fmax(d, v);
  }
  for (int o = 16; o; o >>= 1) {
    a = fmin(a, __shfl_xor_sync(0xffffffffu, a, o)), b = fmax(b, __shfl_xor_sync(0xffffffffu, b, o));
    c = fmin(c, __shfl_xor_sync(0xffffffffu, c, o)), d = fmax(d, __shfl_xor_sync(0xffffffffu, d, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomic_min_d(&mm64[4 * n], a), atomic_max_d(&mm64[4 * n + 1], b), atomic_min_d(&mm64[4 * n + 2], c), atomic_max_d(&mm64[4 * n + 3], d);
  }
}
// cv2.normalize(src32f, NORM_MINMAX, 0, 1, CV_32F): scale = (float)(1/(max-min)), shift = 0.f - (float)(min*scale); dst = src*scale + shift
__global__ void pp_normalize_hv(const float* __restrict__ hv, int P, long long total, const float* __restrict__ mm, float* __restrict__ hn,
                                float* __restrict__ vn) {
  PP_GRID_LOOP(i, total) {
    const int n = (int)(i / P);
    const double hmin = mm[4 * n], hmax = mm[4 * n + 1], vmin = mm[4 * n + 2], vmax = mm[4 * n + 3];
    const float hs = (float)((hmax - hmin > DBL_EPSILON) ? 1.0 / (hmax - hmin) : 0.0);
    const float vs = (float)((vmax - vmin > DBL_EPSILON) ? 1.0 / (vmax - vmin) : 0.0);
    const float hsh = 0.f - (float)(hmin * (double)hs), vsh = 0.f - (float)(vmin * (double)vs);
    // OpenCV's 32F->32F convertTo rounds src*scale + shift once (fused); verified bit-exact against cv2 4.13
    hn[i] = fmaf(hv[2 * i], hs, hsh);
    vn[i] = fmaf(hv[2 * i + 1], vs, vsh);
  }
}
// Sobel rows: rowh = deriv along x of h, rowv = smooth along x of v   (generic RowFilter: s = k[0]*S[0]; s += k[i]*S[i])
template <int K>
__global__ void pp_sobel_rows(const float* __restrict__ hn, const float* __restrict__ vn, int H, int W, long long total,
                              double* __restrict__ rowh, double* __restrict__ rowv) {
  const double* kd = sobel_d<K>();
  const double* ks = sobel_s<K>();
  PP_GRID_LOOP(i, total) {
    const int x = (int)(i % W);
    const long long rowbase = i - x;
    double sh = 0.0, sv = 0.0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const int xx = refl101(x + k - K / 2, W);
      const double a = (double)hn[rowbase + xx], b = (double)vn[rowbase + xx];
      if (k == 0) {
        sh = __dmul_rn(kd[0], a);
        sv = __dmul_rn(ks[0], b);
      } else {
        sh = __dadd_rn(sh, __dmul_rn(kd[k], a));
        sv = __dadd_rn(sv, __dmul_rn(ks[k], b));
      }
    }
    rowh[i] = sh;
    rowv[i] = sv;
  }
}
// Sobel columns: sobel_h = smooth along y of rowh (symmetric), sobel_v = deriv along y of rowv (anti-symmetric)
template <int K>
__global__ void pp_sobel_cols(const double* __restrict__ rowh, const double* __restrict__ rowv, int H, int W, long long total,
                              double* __restrict__ sobh, double* __restrict__ sobv) {
  const double* kd = sobel_d<K>();
  const double* ks = sobel_s<K>();
  constexpr int C = K / 2;
  const int P = H * W;
  PP_GRID_LOOP(i, total) {
    const int p = (int)(i % P);
    const long long base = i - p;
    const int x = p % W, y = p / W;
    double sh = __dmul_rn(ks[C], rowh[i]);
    double sv = 0.0;
#pragma unroll
    for (int k = 1; k <= C; ++k) {
      const int ya = refl101(y + k, H), yb = refl101(y - k, H);
      const double ha = rowh[base + (long long)ya * W + x], hb = rowh[base + (long long)yb * W + x];
      const double va = rowv[base + (long long)ya * W + x], vb = rowv[base + (long long)yb * W + x];
      sh = __dadd_rn(sh, __dmul_rn(ks[C + k], __dadd_rn(ha, hb)));
      sv = __dadd_rn(sv, __dmul_rn(kd[C + k], __dadd_rn(va, -vb)));
    }
    sobh[i] = sh;
    sobv[i] = sv;
  }
}
// 1 - normalize(sobel) (float32), overall, dist0, marker seeds
__global__ void pp_energy(const double* __restrict__ sobh, const double* __restrict__ sobv, const double* __restrict__ mm64,
                          const uint8_t* __restrict__ blb, int P, long long total, double* __restrict__ dist0, uint8_t* __restrict__ mk) {
  PP_GRID_LOOP(i, total) {
    const int n = (int)(i / P);
    const double hmin = mm64[4 * n], hmax = mm64[4 * n + 1], vmin = mm64[4 * n + 2], vmax = mm64[4 * n + 3];
    const double hs = (hmax - hmin > DBL_EPSILON) ? 1.0 / (hmax - hmin) : 0.0;
    const double vs = (vmax - vmin > DBL_EPSILON) ? 1.0 / (vmax - vmin) : 0.0;
    // rtype == CV_32F: scale = (float)scale; shift = (float)dmin - (float)(smin*scale)
    const float hsf = (float)hs, vsf = (float)vs;
    const float hshift = 0.f - (float)(hmin * (double)hsf), vshift = 0.f - (float)(vmin * (double)vsf);
    // convertTo 64F -> 32F: saturate_cast<float>(src * (double)scale + (double)shift)
    const float hn = (float)__dadd_rn(__dmul_rn(sobh[i], (double)hsf), (double)hshift);
    const float vn = (float)__dadd_rn(__dmul_rn(sobv[i], (double)vsf), (double)vshift);
    const float sh = __fadd_rn(1.f, -hn), sv = __fadd_rn(1.f, -vn);
    const float mx = fmaxf(sh, sv);
    const int b = blb[i];
    double overall = (double)mx - (double)(1 - b);
    if (overall < 0.0) overall = 0.0;
    dist0[i] = __dmul_rn(__dadd_rn(1.0, -overall), (double)b);
    const int ov = overall >= 0.4 ? 1 : 0;
    const int m = b - ov;
    mk[i] = m > 0 ? 1 : 0;
  }
}
// dist = -GaussianBlur(dist0, (3,3), 0) in float64, REFLECT_101: rows then columns, out = c*k0 + (l + r)*k1
__global__ void pp_blur_rows(const double* __restrict__ src, int W, long long total, double* __restrict__ dst) {
  PP_GRID_LOOP(i, total) {
    const int x = (int)(i % W);
    const long long rb = i - x;
    const double l = src[rb + refl101(x - 1, W)], r = src[rb + refl101(x + 1, W)];
    // generic row filter: ((k0*l + k1*c) + k2*r)  (bit-exact against cv2.GaussianBlur on float64)
    dst[i] = __dadd_rn(__dadd_rn(__dmul_rn(l, 0.25), __dmul_rn(src[i], 0.5)), __dmul_rn(r, 0.25));
  }
}
__global__ void pp_blur_cols(const double* __restrict__ src, int H, int W, long long total, double* __restrict__ dst) {
  const int P = H * W;
  PP_GRID_LOOP(i, total) {
    const int p = (int)(i % P);
    const long long base = i - p;
    const int x = p % W, y = p / W;
    const double u = src[base + (long long)refl101(y - 1, H) * W + x], d = src[base + (long long)refl101(y + 1, H) * W + x];
    dst[i] = -__dadd_rn(__dmul_rn(src[i], 0.5), __dmul_rn(__dadd_rn(u, d), 0.25));
  }
}
__global__ void pp_invert(const uint8_t* __restrict__ a, long long total, uint8_t* __restrict__ b) {
  PP_GRID_LOOP(i, total) b[i] = a[i] ? 0 : 1;
}
// filled = marker | (background component that does not touch the border)
__global__ void pp_fill_holes(const uint8_t* __restrict__ mk, const int* __restrict__ lab, const uint8_t* __restrict__ border, int P,
                              long long total, uint8_t* __restrict__ out) {
  PP_GRID_LOOP(i, total) {
    const long long base = i - (i % P);
    const int r = lab[i];
    out[i] = (mk[i] || (r >= 0 && !border[base + r])) ? 1 : 0;
  }
}
// cv2 MORPH_ELLIPSE (5,5): rows 00100 / 11111 / 11111 / 11111 / 00100 ; erosion ignores out-of-image, dilation treats it as 0
__device__ __forceinline__ bool in_ellipse5(int dy, int dx) { return (dy == -2 || dy == 2) ? dx == 0 : true; }
__global__ void pp_morph(const uint8_t* __restrict__ src, int H, int W, long long total, int dilate, uint8_t* __restrict__ dst) {
  const int P = H * W;
  PP_GRID_LOOP(i, total) {
    const int p = (int)(i % P);
    const long long base = i - p;
    const int x = p % W, y = p / W;
    int acc = dilate ? 0 : 1;
    for (int dy = -2; dy <= 2; ++dy)
      for (int dx = -2; dx <= 2; ++dx) {
        if (!in_ellipse5(dy, dx)) continue;
        const int yy = y + dy, xx = x + dx;
        if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
        const int v = src[base + (long long)yy * W + xx];
        acc = dilate ? (acc | v) : (acc & v);
      }
    dst[i] = (uint8_t)acc;
  }
}

// heap slice offsets of the foreground components: one thread per image (components are few)
// Exclusive prefix sum of one int per thread over a PP_THREADS block; *total = block sum.  `sh` holds >= 33 ints.
__device__ __forceinline__ int block_excl_scan(int v, int* sh, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += t;
  }
  __syncthreads();   // sh may still be read by the previous call
  if (lane == 31) sh[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = lane < (int)(blockDim.x >> 5) ? sh[lane] : 0;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, w, d);
      if (lane >= d) w += t;
    }
    sh[lane] = w;   // inclusive warp totals
  }
  __syncthreads();
  *total = sh[(blockDim.x >> 5) - 1];
  return inc - v + (warp ? sh[warp - 1] : 0);
}

// heap arena offset of every component = exclusive prefix sum of the component sizes in rank order; one block per image
__global__ void __launch_bounds__(PP_THREADS) pp_heap_offsets(const int* __restrict__ roots, const int* __restrict__ ncomp,
                                                              const int* __restrict__ size, int P, int N, int* __restrict__ heap_off) {
  __shared__ int sh[33];
  for (int n = blockIdx.x; n < N; n += gridDim.x) {
    const long long base = (long long)n * P;
    const int nc = ncomp[n];
    int run = 0;
    for (int c0 = 0; c0 < nc; c0 += blockDim.x) {
      const int c = c0 + threadIdx.x;
      int r = -1, v = 0;
      if (c < nc) {
        r = roots[base + c];
        v = size[base + r];
      }
      int tot;
      const int ex = block_excl_scan(v, sh, &tot);
      if (r >= 0) heap_off[base + r] = run + ex;
      run += tot;
    }
  }
}

__device__ __forceinline__ bool heap_less(const HeapItem& a, const HeapItem& b) {
  return a.v != b.v ? a.v < b.v : a.age < b.age;
}
__device__ __forceinline__ void heap_push(HeapItem* h, int& n, HeapItem e) {
  int c = n++;
  while (c > 0) {
    const int par = (c - 1) >> 1;
    if (heap_less(e, h[par])) {
      h[c] = h[par];
      c = par;
    } else {
      break;
    }
  }
  h[c] = e;
}
__device__ __forceinline__ HeapItem heap_pop(HeapItem* h, int& n) {
  const HeapItem top = h[0];
  const HeapItem last = h[--n];
  int p = 0;
  while (true) {
    int c = 2 * p + 1;
    if (c >= n) break;
    if (c + 1 < n && heap_less(h[c + 1], h[c])) ++c;
    if (heap_less(h[c], last)) {
      h[p] = h[c];
      p = c;
    } else {
      break;
    }
  }
  if (n > 0) h[p] = last;
  return top;
}
// marker-controlled watershed, one thread per (image, foreground component)
__global__ void __launch_bounds__(64) pp_watershed(const int* __restrict__ roots, const int* __restrict__ ncomp, const int* __restrict__ lab,
                                                   const int* __restrict__ size, const int* __restrict__ bbox, const int* __restrict__ heap_off,
                                                   const uint8_t* __restrict__ blb, const int* __restrict__ marker,
                                                   const double* __restrict__ dist, int H, int W, int N, int maxcomp,
                                                   HeapItem* __restrict__ heap, int* __restrict__ inst) {
  const int P = H * W;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int n = (int)(t / maxcomp), c = (int)(t % maxcomp);
  if (n >= N || c >= ncomp[n]) return;
  const long long base = (long long)n * P;
  const int r = roots[base + c];
  if (size[base + r] <= 9) return;  // removed small object: not part of the mask
  HeapItem* hp = heap + base + heap_off[base + r];
  int hn = 0;
  const int* bb = bbox + 4 * (base + r);
  // markers of this component, in raveled order, all with age 0
  for (int y = bb[1]; y <= bb[3]; ++y)
    for (int x = bb[0]; x <= bb[2]; ++x) {
      const int p = y * W + x;
      if (lab[base + p] != r) continue;
      const int m = marker[base + p];
      if (m > 0) {
        inst[base + p] = m;
        heap_push(hp, hn, HeapItem{dist[base + p], 0u, (unsigned)p});
      }
    }
  unsigned age = 0;
  while (hn > 0) {
    const HeapItem e = heap_pop(hp, hn);
    const int p = (int)e.idx;
    const int x = p % W, y = p / W;
    const int l = inst[base + p];
    const int nb[4] = {p - W, p - 1, p + 1, p + W};
    const bool ok[4] = {y > 0, x > 0, x < W - 1, y < H - 1};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (!ok[k]) continue;
      const int q = nb[k];
      if (!blb[base + q] || inst[base + q]) continue;
      ++age;
      inst[base + q] = l;
      heap_push(hp, hn, HeapItem{dist[base + q], age, (unsigned)q});
    }
  }
}

// ---- instance table ----
__global__ void pp_inst_init(long long totalI, int* iarea, int* ibox, unsigned long long* isx, unsigned long long* isy, int* itype) {
  PP_GRID_LOOP(i, totalI) {
    iarea[i] = 0;
    ibox[4 * i] = 0x7fffffff, ibox[4 * i + 1] = 0x7fffffff, ibox[4 * i + 2] = -1, ibox[4 * i + 3] = -1;
    isx[i] = 0ull, isy[i] = 0ull;
    for (int k = 0; k < PP_MAX_TYPES; ++k) itype[i * PP_MAX_TYPES + k] = 0;
  }
}
__global__ void pp_inst_accum(const int* __restrict__ inst, const float* __restrict__ tp, int H, int W, long long total, int maxI,
                              int* iarea, int* ibox, unsigned long long* isx, unsigned long long* isy, int* itype) {
  const int P = H * W;
  PP_GRID_LOOP(i, total) {
    const int id = inst[i];
    if (id <= 0 || id > maxI) continue;
    const int n = (int)(i / P), p = (int)(i % P);
    const int x = p % W, y = p / W;
    const long long s = (long long)n * maxI + (id - 1);
    atomicAdd(&iarea[s], 1);
    atomicMin(&ibox[4 * s], x), atomicMin(&ibox[4 * s + 1], y), atomicMax(&ibox[4 * s + 2], x), atomicMax(&ibox[4 * s + 3], y);
    atomicAdd(&isx[s], (unsigned long long)x), atomicAdd(&isy[s], (unsigned long long)y);
    if (tp) {
      int t = (int)(uint8_t)rintf(tp[i]);  // np.around(tp).astype("uint8")
      if (t >= PP_MAX_TYPES) t = PP_MAX_TYPES - 1;
      atomicAdd(&itype[s * PP_MAX_TYPES + t], 1);
    }
  }
}
// compact rows in ascending id order: one block per image, block-wide scan of the "instance exists" flags
__global__ void __launch_bounds__(PP_THREADS) pp_inst_finalize(int N, int maxI, int has_type, const int* __restrict__ iarea,
                                                               const int* __restrict__ ibox, const unsigned long long* __restrict__ isx,
                                                               const unsigned long long* __restrict__ isy, const int* __restrict__ itype,
                                                               int* __restrict__ count, int* __restrict__ ids, int* __restrict__ box,
                                                               double* __restrict__ centroid, int* __restrict__ type,
                                                               double* __restrict__ prob) {
  __shared__ int sh[33];
  for (int n = blockIdx.x; n < N; n += gridDim.x) {
    int run = 0;
    for (int i0 = 0; i0 < maxI; i0 += blockDim.x) {
      const int id = i0 + threadIdx.x + 1;
      const long long s = (long long)n * maxI + (id - 1);
      const int a = id <= maxI ? iarea[s] : 0;
      int tot;
      const int k = run + block_excl_scan(a > 0 ? 1 : 0, sh, &tot);
      run += tot;
      if (a <= 0) continue;
      const long long o = (long long)n * maxI + k;
      ids[o] = id;
      box[4 * o] = ibox[4 * s], box[4 * o + 1] = ibox[4 * s + 1], box[4 * o + 2] = ibox[4 * s + 2] + 1, box[4 * o + 3] = ibox[4 * s + 3] + 1;
      // cv2.moments on the crop: m10/m00 + box origin == mean absolute coordinate
      centroid[2 * o] = (double)(isx[s] - (unsigned long long)ibox[4 * s] * a) / (double)a + (double)ibox[4 * s];
      centroid[2 * o + 1] = (double)(isy[s] - (unsigned long long)ibox[4 * s + 1] * a) / (double)a + (double)ibox[4 * s + 1];
      if (has_type) {
        // sorted by count descending, stable -> ties keep ascending type id
        int best = -1, bc = -1, second = -1, sc = -1;
        for (int t = 0; t < PP_MAX_TYPES; ++t) {
          const int c = itype[s * PP_MAX_TYPES + t];
          if (c <= 0) continue;
          if (c > bc) {
            second = best, sc = bc;
            best = t, bc = c;
          } else if (c > sc) {
            second = t, sc = c;
          }
        }
        int ty = best, tc = bc;
        if (best == 0 && second >= 0) ty = second, tc = sc;
        type[o] = ty;
        prob[o] = (double)tc / ((double)a + 1.0e-6);
      } else {
        type[o] = -1;
        prob[o] = 0.0;
      }
    }
    if (threadIdx.x == 0) count[n] = run;
  }
}

// Version 2 of the flood: one WARP per component (lane 0 walks, so components never wait on each other inside a divergent warp) over the
// 8-ary heap of watershed_math.cuh, total order (value, age, index).  Selected with tiab_hovernet_set_watershed(h, 2).
__global__ void __launch_bounds__(128) pp_watershed_v2(const int* __restrict__ roots, const int* __restrict__ ncomp, const int* __restrict__ lab,
                                                       const int* __restrict__ size, const int* __restrict__ bbox, const int* __restrict__ heap_off,
                                                       const uint8_t* __restrict__ blb, const int* __restrict__ marker,
                                                       const double* __restrict__ dist, int H, int W, int N, int maxcomp,
                                                       HeapItem* __restrict__ heap, int* __restrict__ inst) {
  if (threadIdx.x & 31) return;
  const int P = H * W;
  const long long t = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int n = (int)(t / maxcomp), c = (int)(t % maxcomp);
  if (n >= N || c >= ncomp[n]) return;
  const long long base = (long long)n * P;
  const int r = roots[base + c];
  if (size[base + r] <= 9) return;
  const int* bb = bbox + 4 * (base + r);
  static_assert(sizeof(HeapItem) == sizeof(WsItem), "the heap arena is shared between the two flood versions");
  ws_flood<8>(lab + base, r, bb[0], bb[1], bb[2], bb[3], blb + base, marker + base, dist + base, H, W,
              reinterpret_cast<WsItem*>(heap + base + heap_off[base + r]), inst + base);
}

// Version 3 of the flood: the whole working set of a component in SHARED memory.  r1h measured ~9 500 cycles per flooded pixel for
// version 1 — every heap level and every neighbour test was an L2 round trip.  Here a persistent single-warp block takes one live component at
// a time from a work list, stages the component's bounding box (membership, markers, distances) and its heap in shared memory (13 B per
// box pixel + 16 B per component pixel), lane 0 runs the same ws_flood<8> in box-local coordinates (row-major like the image, so the
// index tie-break is unchanged), and the warp writes the labels back.  Components that do not fit fall back to the global-memory flood.
__global__ void pp_flood_worklist(const int* __restrict__ roots, const int* __restrict__ ncomp, const int* __restrict__ size, int P, int N,
                                  int maxcomp, int size_lo, int size_hi, int* __restrict__ wlist, int* __restrict__ wcount) {
  // called once per size class [size_lo, size_hi), largest class first: the persistent flood blocks then start with the long poles
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)N * maxcomp) return;
  const int n = (int)(t / maxcomp), c = (int)(t % maxcomp);
  if (c >= ncomp[n]) return;
  const int r = roots[(long long)n * P + c];
  const int sz = size[(long long)n * P + r];
  if (sz <= 9 || sz < size_lo || sz >= size_hi) return;
  const int slot = atomicAdd(wcount, 1);
  wlist[2 * slot] = n;
  wlist[2 * slot + 1] = r;
}

__global__ void __launch_bounds__(32) pp_watershed_v3(const int* __restrict__ wlist, int* __restrict__ wcount, const int* __restrict__ lab,
                                                      const int* __restrict__ size, const int* __restrict__ bbox,
                                                      const int* __restrict__ heap_off, const uint8_t* __restrict__ blb,
                                                      const int* __restrict__ marker, const double* __restrict__ dist, int H, int W,
                                                      HeapItem* __restrict__ heap, int* __restrict__ inst, int smem_bytes) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int P = H * W;
  const int lane = threadIdx.x;
  const int total = wcount[0];
  for (;;) {
    int item = 0;
    if (lane == 0) item = atomicAdd(wcount + 1, 1);
    item = __shfl_sync(0xffffffffu, item, 0);
    if (item >= total) break;
    const int n = wlist[2 * item], r = wlist[2 * item + 1];
    const long long base = (long long)n * P;
    const int* bb = bbox + 4 * (base + r);
    const int x0 = bb[0], y0 = bb[1], bw = bb[2] - bb[0] + 1, bh = bb[3] - bb[1] + 1;
    const int area = bw * bh, npx = size[base + r];
    // shared-memory carve: heap | dist (8 B) | labels (4 B) | membership (1 B)
    const size_t need = (size_t)npx * sizeof(WsItem) + (size_t)area * 13 + 32;
    if (need > (size_t)smem_bytes) {   // too large: flood in place (global memory), same code
      if (lane == 0)
        ws_flood<8>(lab + base, r, bb[0], bb[1], bb[2], bb[3], blb + base, marker + base, dist + base, H, W,
                    reinterpret_cast<WsItem*>(heap + base + heap_off[base + r]), inst + base);
      __syncwarp();
      continue;
    }
    WsItem* hp = reinterpret_cast<WsItem*>(smem);
    double* ldist = reinterpret_cast<double*>(smem + (((size_t)npx * sizeof(WsItem) + 15) & ~(size_t)15));
    int* lout = reinterpret_cast<int*>(ldist + area);
    uint8_t* lmask = reinterpret_cast<uint8_t*>(lout + area);
    for (int q = lane; q < area; q += 32) {
      const int ly = q / bw, lx = q - ly * bw;
      const long long g = base + (long long)(y0 + ly) * W + (x0 + lx);
      const bool in = lab[g] == r;
      lmask[q] = in ? 1 : 0;
      lout[q] = in ? marker[g] : 0;      // markers of this component seed the flood; ws_flood labels in place
      ldist[q] = dist[g];
    }
    __syncwarp();
    if (lane == 0) ws_flood<8>(nullptr, 0, 0, 0, bw - 1, bh - 1, lmask, lout, ldist, bh, bw, hp, lout);
    __syncwarp();
    for (int q = lane; q < area; q += 32)
      if (lmask[q]) {
        const int ly = q / bw, lx = q - ly * bw;
        inst[base + (long long)(y0 + ly) * W + (x0 + lx)] = lout[q];
      }
    __syncwarp();
  }
}

// Version 4 of the flood: the WHOLE WARP works on one component.  r2b measured version 3 at ~2 900 cycles per flooded pixel on its
// longest component: lane 0 walks an 8-ary heap in shared memory and every level is a chain of dependent LDS + DSETP + branch.  Here the
// priority queue is a 32-ary implicit heap in shared memory (struct-of-arrays, 32-bit words): a node's 32 children are loaded by the 32
// lanes in ONE shared-memory round trip and their minimum is found with `redux.sync.min` on the high word of an order-preserving 64-bit
// image of the distance (a ballot tells whether the low word / tie-break word are needed at all), so a pop costs two levels for queues of
// up to 1 057 items.  The total order is the same as everywhere else — (value, age, raveled index), folded into one tie-break word
// `t` = index for markers (age 0) and 0x80000000 | age otherwise — so any heap shape gives the oracle's labels.  The component's
// bounding box is staged with a one-pixel frame (label -1 = not in the component) so that the four neighbour tests need no bounds
// checks and are made by four lanes at once.  Components whose working set exceeds the shared-memory budget use the version-2 flood.
struct WHeap {
  unsigned long long* key;   // order-preserving 64-bit image of the distance
  unsigned long long* ti;    // (tie-break word << 32) | local pixel index
  int n;
};
__device__ __forceinline__ bool wkey_less(unsigned long long ak, uint32_t at, unsigned long long bk, uint32_t bt) {
  return ak != bk ? ak < bk : at < bt;
}
// order-preserving map double -> uint64 (no NaNs; -0.0 canonicalised to +0.0 so that equal doubles give equal keys)
__device__ __forceinline__ unsigned long long wkey_of(double d) {
  const long long b = __double_as_longlong(d + 0.0);
  return b < 0 ? ~(unsigned long long)b : ((unsigned long long)b | 0x8000000000000000ull);
}
// minimum of the lanes' items under the total order (key, tie-break word): returns the winning lane, its item in (wk, wti).
// One redux.sync on the key's high word settles it unless several lanes share that word (plateaus / exact ties).
__device__ __forceinline__ int warp_argmin(unsigned long long ck, unsigned long long cti, bool valid, unsigned long long& wk,
                                           unsigned long long& wti) {
  const unsigned FULL = 0xffffffffu;
  const uint32_t ch = (uint32_t)(ck >> 32), cl = (uint32_t)ck, ct = (uint32_t)(cti >> 32);
  const uint32_t r1 = __reduce_min_sync(FULL, valid ? ch : 0xffffffffu);
  unsigned e = __ballot_sync(FULL, valid && ch == r1);
  if (__popc(e) != 1) {
    const bool m1 = valid && ch == r1;
    const uint32_t r2 = __reduce_min_sync(FULL, m1 ? cl : 0xffffffffu);
    e = __ballot_sync(FULL, m1 && cl == r2);
    if (__popc(e) != 1) {
      const bool m2 = m1 && cl == r2;
      const uint32_t r3 = __reduce_min_sync(FULL, m2 ? ct : 0xffffffffu);
      e = __ballot_sync(FULL, m2 && ct == r3);
    }
  }
  const int w = __ffs(e) - 1;
  wk = __shfl_sync(FULL, ck, w), wti = __shfl_sync(FULL, cti, w);
  return w;
}
// place item (ik, iti) at node p of the shared-memory heap and sift it down; every lane passes the same arguments
__device__ __forceinline__ void wheap_sift_down(WHeap& h, int p, unsigned long long ik, unsigned long long iti, int lane) {
  const int n = h.n;
  const uint32_t it = (uint32_t)(iti >> 32);
  for (;;) {
    const int c0 = 32 * p + 1;
    if (c0 >= n) break;
    const int c = c0 + lane;
    const bool valid = c < n;
    const unsigned long long ck = valid ? h.key[c] : ~0ull;
    const unsigned long long cti = valid ? h.ti[c] : ~0ull;
    unsigned long long wk, wti;
    const int w = warp_argmin(ck, cti, valid, wk, wti);
    if (!wkey_less(wk, (uint32_t)(wti >> 32), ik, it)) break;
    if (lane == 0) h.key[p] = wk, h.ti[p] = wti;
    p = c0 + w;
  }
  if (lane == 0) h.key[p] = ik, h.ti[p] = iti;
  __syncwarp();
}
// The two top levels of the heap live in REGISTERS during the flood: the root in every lane (uniform), level-1 node j (heap index 1 + j)
// in lane j.  A sift-down from the root then starts with a register-only argmin and touches shared memory only when the item sinks
// below level 1; `n` (h.n) counts all items as before, heap indices >= 33 stay in shared memory.
struct WTop {
  unsigned long long rk, rti, lk, lti;
};
__device__ __forceinline__ void wtop_sift_root(WHeap& h, WTop& t, unsigned long long ik, unsigned long long iti, int lane) {
  const uint32_t it = (uint32_t)(iti >> 32);
  const int n = h.n;
  if (n <= 1) {
    t.rk = ik, t.rti = iti;
    return;
  }
  unsigned long long wk, wti;
  int w = warp_argmin(t.lk, t.lti, 1 + lane < n, wk, wti);
  if (!wkey_less(wk, (uint32_t)(wti >> 32), ik, it)) {
    t.rk = ik, t.rti = iti;
    return;
  }
  t.rk = wk, t.rti = wti;
  int p = 1 + w;          // the item sinks into level-1 node w (a register of lane w) or below it (shared memory)
  const int reg_lane = w;
  bool in_regs = true;
  for (;;) {
    const int c0 = 32 * p + 1;
    if (c0 >= n) break;
    const int c = c0 + lane;
    const bool valid = c < n;
    const unsigned long long ck = valid ? h.key[c] : ~0ull;
    const unsigned long long cti = valid ? h.ti[c] : ~0ull;
    w = warp_argmin(ck, cti, valid, wk, wti);
    if (!wkey_less(wk, (uint32_t)(wti >> 32), ik, it)) break;
    if (in_regs) {
      if (lane == reg_lane) t.lk = wk, t.lti = wti;
      in_regs = false;
    } else if (lane == 0) {
      h.key[p] = wk, h.ti[p] = wti;
    }
    p = c0 + w;
  }
  if (in_regs) {
    if (lane == reg_lane) t.lk = ik, t.lti = iti;
  } else {
    if (lane == 0) h.key[p] = ik, h.ti[p] = iti;
    __syncwarp();
  }
}
__device__ __forceinline__ void wtop_push(WHeap& h, WTop& t, unsigned long long ik, unsigned long long iti, int lane) {
  const unsigned FULL = 0xffffffffu;
  const uint32_t it = (uint32_t)(iti >> 32);
  int c = h.n++;
  if (c == 0) {
    t.rk = ik, t.rti = iti;
    return;
  }
  bool wrote = false;
  while (c >= 33) {
    const int par = (c - 1) >> 5;
    unsigned long long pk, pti;
    if (par >= 33) {
      pk = h.key[par], pti = h.ti[par];
    } else {
      pk = __shfl_sync(FULL, t.lk, par - 1), pti = __shfl_sync(FULL, t.lti, par - 1);
    }
    if (!wkey_less(ik, it, pk, (uint32_t)(pti >> 32))) break;
    if (lane == 0) h.key[c] = pk, h.ti[c] = pti;
    wrote = true;
    c = par;
  }
  if (c >= 33) {
    if (lane == 0) h.key[c] = ik, h.ti[c] = iti;
    wrote = true;
  } else if (wkey_less(ik, it, t.rk, (uint32_t)(t.rti >> 32))) {   // level 1 (lane c - 1), then against the root
    if (lane == c - 1) t.lk = t.rk, t.lti = t.rti;
    t.rk = ik, t.rti = iti;
  } else if (lane == c - 1) {
    t.lk = ik, t.lti = iti;
  }
  if (wrote) __syncwarp();
}
__device__ __forceinline__ void wheap_push(WHeap& h, unsigned long long ik, unsigned long long iti, int lane) {
  const uint32_t it = (uint32_t)(iti >> 32);
  int c = h.n++;
  while (c > 0) {
    const int par = (c - 1) >> 5;
    const unsigned long long pk = h.key[par], pti = h.ti[par];
    if (!wkey_less(ik, it, pk, (uint32_t)(pti >> 32))) break;
    if (lane == 0) h.key[c] = pk, h.ti[c] = pti;
    c = par;
  }
  if (lane == 0) h.key[c] = ik, h.ti[c] = iti;
  __syncwarp();
}

__global__ void __launch_bounds__(32) pp_watershed_v4(const int* __restrict__ wlist, int* __restrict__ wcount, const int* __restrict__ lab,
                                                      const int* __restrict__ size, const int* __restrict__ bbox,
                                                      const int* __restrict__ heap_off, const uint8_t* __restrict__ blb,
                                                      const int* __restrict__ marker, const double* __restrict__ dist, int H, int W,
                                                      HeapItem* __restrict__ heap, int* __restrict__ inst, int smem_bytes) {
  extern __shared__ __align__(16) unsigned char smem[];
  const unsigned FULL = 0xffffffffu;
  const int P = H * W;
  const int lane = threadIdx.x;
  const int total = wcount[0];
  for (;;) {
    int item = 0;
    if (lane == 0) item = atomicAdd(wcount + 1, 1);
    item = __shfl_sync(FULL, item, 0);
    if (item >= total) break;
    const int n_img = wlist[2 * item], r = wlist[2 * item + 1];
    const long long base = (long long)n_img * P;
    const int* bb = bbox + 4 * (base + r);
    const int x0 = bb[0], y0 = bb[1], bw = bb[2] - bb[0] + 1, bh = bb[3] - bb[1] + 1;
    const int pw = bw + 2, ph = bh + 2, parea = pw * ph, npx = size[base + r];
    const int parea2 = (parea + 1) & ~1;
    const size_t need = (size_t)parea2 * 12 + (size_t)npx * 16 + 64;
    if (need > (size_t)smem_bytes) {   // too large for shared memory: flood in place (global memory), lane 0, 8-ary heap
      if (lane == 0)
        ws_flood<8>(lab + base, r, bb[0], bb[1], bb[2], bb[3], blb + base, marker + base, dist + base, H, W,
                    reinterpret_cast<WsItem*>(heap + base + heap_off[base + r]), inst + base);
      __syncwarp();
      continue;
    }
    unsigned long long* pkey = reinterpret_cast<unsigned long long*>(smem);
    WHeap h;
    h.key = pkey + parea2;
    h.ti = h.key + npx;
    int* lout = reinterpret_cast<int*>(h.ti + npx);
    h.n = 0;
    // stage the framed box; marker pixels go straight into the (not yet ordered) heap array, tie-break word = raveled local index
    for (int q0 = 0; q0 < parea; q0 += 32) {
      const int q = q0 + lane;
      bool is_marker = false;
      unsigned long long k = 0ull;
      if (q < parea) {
        const int ly = q / pw, lx = q - ly * pw;
        int lo = -1;
        if (lx >= 1 && lx <= bw && ly >= 1 && ly <= bh) {
          const long long g = base + (long long)(y0 + ly - 1) * W + (x0 + lx - 1);
          if (lab[g] == r) {
            lo = marker[g];
            k = wkey_of(dist[g]);
            is_marker = lo > 0;
          }
        }
        lout[q] = lo;
        pkey[q] = k;
      }
      const unsigned mk = __ballot_sync(FULL, is_marker);
      if (is_marker) {
        const int pos = h.n + __popc(mk & ((1u << lane) - 1u));
        h.key[pos] = k;
        h.ti[pos] = ((unsigned long long)(uint32_t)q << 32) | (uint32_t)q;
      }
      h.n += __popc(mk);
    }
    __syncwarp();
    // heapify (Floyd): sift every internal node down, last parent first
    for (int p = (h.n - 2) >> 5; p >= 0 && h.n > 1; --p) {
      const unsigned long long ik = h.key[p], iti = h.ti[p];
      __syncwarp();
      wheap_sift_down(h, p, ik, iti, lane);
    }
    // the flood.  The popped pixel's first new neighbour REPLACES the root (one sift-down instead of pop + push); only when there is
    // none does the last item move to the root.  Root and level 1 of the heap are held in registers (WTop).
    WTop top;
    top.rk = top.rti = top.lk = top.lti = ~0ull;
    if (h.n > 0) top.rk = h.key[0], top.rti = h.ti[0];
    if (1 + lane < h.n) top.lk = h.key[1 + lane], top.lti = h.ti[1 + lane];
    uint32_t age = 0;
    const int off = lane == 0 ? -pw : (lane == 1 ? -1 : (lane == 2 ? 1 : pw));
    while (h.n > 0) {
      const int pidx = (int)(uint32_t)top.rti;
      const int label = lout[pidx];
      const int q = pidx + off;
      const bool cand = lane < 4 && lout[q] == 0;
      unsigned bits = __ballot_sync(FULL, cand);
      if (cand) lout[q] = label;
      if (bits) {
        const int k = __ffs(bits) - 1;
        bits &= bits - 1;
        const int qq = pidx + (k == 0 ? -pw : (k == 1 ? -1 : (k == 2 ? 1 : pw)));
        ++age;
        wtop_sift_root(h, top, pkey[qq], ((unsigned long long)(0x80000000u | age) << 32) | (uint32_t)qq, lane);
      } else {
        const int last = --h.n;
        if (last > 0) {
          unsigned long long lk, lti;
          if (last >= 33) {
            lk = h.key[last], lti = h.ti[last];
          } else {
            lk = __shfl_sync(FULL, top.lk, last - 1), lti = __shfl_sync(FULL, top.lti, last - 1);
          }
          wtop_sift_root(h, top, lk, lti, lane);
        }
      }
      while (bits) {
        const int k = __ffs(bits) - 1;
        bits &= bits - 1;
        const int qq = pidx + (k == 0 ? -pw : (k == 1 ? -1 : (k == 2 ? 1 : pw)));
        ++age;
        wtop_push(h, top, pkey[qq], ((unsigned long long)(0x80000000u | age) << 32) | (uint32_t)qq, lane);
      }
      __syncwarp();
    }
    // labels back to the instance map (members only; unreached members stay 0 like in the reference)
    for (int q = lane; q < parea; q += 32) {
      const int v = lout[q];
      if (v > 0) {
        const int ly = q / pw, lx = q - ly * pw;
        inst[base + (long long)(y0 + ly - 1) * W + (x0 + lx - 1)] = v;
      }
    }
    __syncwarp();
  }
}

// Contours of the tabulated instances (hovernet.py:685-693): one thread per table row walks the instance's outer border twice — once to
// count the CHAIN_APPROX_SIMPLE vertices, once to write them into a slice reserved with one atomicAdd on the image's vertex counter.
__global__ void __launch_bounds__(128) pp_contours(const int32_t* __restrict__ inst_map, int N, int H, int W, int max_inst,
                                                   const int* __restrict__ inst_count, const int32_t* __restrict__ inst_id,
                                                   const int32_t* __restrict__ inst_box, int capacity, int32_t* __restrict__ cont_off,
                                                   int32_t* __restrict__ cont_len, int32_t* __restrict__ cont_total,
                                                   int32_t* __restrict__ cont_xy) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)N * max_inst) return;
  const int n = (int)(t / max_inst), k = (int)(t - (long long)n * max_inst);
  if (k >= inst_count[n]) return;
  const LabelView v{inst_map + (long long)n * H * W, H, W, inst_id[t]};
  const int32_t* box = inst_box + t * 4;   // [x0, y0, x1, y1): the first pixel in raster order lies on row y0
  int x0 = box[0];
  const int y0 = box[1];
  while (x0 < box[2] && !v.fg(x0, y0)) ++x0;
  if (x0 >= box[2]) {   // cannot happen for a consistent table
    cont_off[t] = 0;
    cont_len[t] = 0;
    return;
  }
  const int len = trace_outer_contour(v, x0, y0, nullptr, 0);
  const int off = atomicAdd(&cont_total[n], len);
  cont_off[t] = off;
  cont_len[t] = len;
  if (off + len <= capacity) trace_outer_contour(v, x0, y0, cont_xy + ((long long)n * capacity + off) * 2, len);
}

}  // namespace tiab

using namespace tiab;

static void sobel_kernels(double* d, double* s, int K = 21) {
  // binomial smoothing (1+x)^(K-1) and derivative (1+x)^(K-2) * (x-1)  (cv2.getDerivKernels(1, 0, K), un-normalised), K <= 21
  long long a[22] = {0}, b[22] = {0};
  a[0] = 1;
  for (int it = 0; it < K - 1; ++it)
    for (int k = it + 1; k >= 1; --k) a[k] += a[k - 1];
  for (int k = 0; k < K; ++k) s[k] = (double)a[k];
  b[0] = 1;
  for (int it = 0; it < K - 2; ++it)
    for (int k = it + 1; k >= 1; --k) b[k] += b[k - 1];
  for (int k = 0; k < K; ++k) d[k] = (double)((k >= 1 ? b[k - 1] : 0) - (k <= K - 2 ? b[k] : 0));
}

struct PPAlloc {
  std::vector<std::pair<void**, size_t>> items;
};

static int pp_ensure(tiab_ctx* h, size_t bytes) {
  if (h->pp_bytes >= bytes) return TIAB_OK;
  if (h->pp_work) cudaFree(h->pp_work);
  h->pp_work = nullptr;
  h->pp_bytes = 0;
  TIAB_CUDA_OK(h, cudaMalloc(&h->pp_work, bytes));
  h->pp_bytes = bytes;
  return TIAB_OK;
}

#define PP_LAUNCH(h, name, grid, block, ...)                                  \
  do {                                                                        \
    name<<<grid, block, 0, s>>>(__VA_ARGS__);                                 \
    cudaError_t _e = cudaGetLastError();                                      \
    if (_e != cudaSuccess) {                                                  \
      (h)->err = std::string(#name " launch: ") + cudaGetErrorString(_e);     \
      return TIAB_E_CUDA;                                                     \
    }                                                                         \
    (h)->last_launches++;                                                     \
  } while (0)

extern "C" {

int tiab_dbg_sobel_kernels(double deriv[21], double smooth[21]) {
  sobel_kernels(deriv, smooth);
  return TIAB_OK;
}

int tiab_hovernet_set_watershed(tiab_handle_t h, int impl) {
  if (!h || impl < 1 || impl > 4) return TIAB_E_INVALID;
  h->pp_watershed_impl = impl;
  return TIAB_OK;
}

int tiab_hovernet_get_watershed(tiab_handle_t h) { return h ? h->pp_watershed_impl : TIAB_E_INVALID; }

int tiab_dbg_watershed(const double* dist_host, const int32_t* markers_host, const uint8_t* mask_host, int H, int W, int32_t* out_host) {
  if (!dist_host || !markers_host || !mask_host || !out_host || H <= 0 || W <= 0) return TIAB_E_INVALID;
  const long long P = (long long)H * W;
  std::vector<WsItem> heap((size_t)P);
  std::vector<int> marker((size_t)P);
  for (long long p = 0; p < P; ++p) marker[p] = mask_host[p] ? markers_host[p] : 0, out_host[p] = 0;   // out = markers * mask
  ws_flood<8>(nullptr, 0, 0, 0, W - 1, H - 1, mask_host, marker.data(), dist_host, H, W, heap.data(), out_host);
  return TIAB_OK;
}

// the version-3 data flow on the host: per 4-connected mask component, box-local copies + ws_flood in local coordinates + write back
int tiab_dbg_watershed_local(const double* dist_host, const int32_t* markers_host, const uint8_t* mask_host, int H, int W, int32_t* out_host) {
  if (!dist_host || !markers_host || !mask_host || !out_host || H <= 0 || W <= 0) return TIAB_E_INVALID;
  const int P = H * W;
  std::vector<int> comp((size_t)P, -1), stack;
  for (int p = 0; p < P; ++p) out_host[p] = 0;
  for (int seed = 0; seed < P; ++seed) {
    if (!mask_host[seed] || comp[seed] >= 0) continue;
    int x0 = W, y0 = H, x1 = -1, y1 = -1, npx = 0;
    stack.assign(1, seed);
    comp[seed] = seed;
    while (!stack.empty()) {
      const int p = stack.back();
      stack.pop_back();
      ++npx;
      const int x = p % W, y = p / W;
      x0 = std::min(x0, x), x1 = std::max(x1, x), y0 = std::min(y0, y), y1 = std::max(y1, y);
      const int nb[4] = {p - W, p - 1, p + 1, p + W};
      const bool ok[4] = {y > 0, x > 0, x < W - 1, y < H - 1};
      for (int k = 0; k < 4; ++k)
        if (ok[k] && mask_host[nb[k]] && comp[nb[k]] < 0) comp[nb[k]] = seed, stack.push_back(nb[k]);
    }
    const int bw = x1 - x0 + 1, bh = y1 - y0 + 1, area = bw * bh;
    std::vector<WsItem> heap((size_t)npx);
    std::vector<double> ldist((size_t)area);
    std::vector<int> lout((size_t)area);
    std::vector<uint8_t> lmask((size_t)area);
    for (int q = 0; q < area; ++q) {
      const int g = (y0 + q / bw) * W + x0 + q % bw;
      const bool in = comp[g] == seed;
      lmask[q] = in, lout[q] = in ? markers_host[g] : 0, ldist[q] = dist_host[g];
    }
    ws_flood<8>(nullptr, 0, 0, 0, bw - 1, bh - 1, lmask.data(), lout.data(), ldist.data(), bh, bw, heap.data(), lout.data());
    for (int q = 0; q < area; ++q)
      if (lmask[q]) out_host[(y0 + q / bw) * W + x0 + q % bw] = lout[q];
  }
  return TIAB_OK;
}

// The work-list floods (versions 3 and 4) on arbitrary host inputs: the whole mask is handed over as ONE work item (bounding box of the
// mask; a disconnected mask is fine for the flood), so that the adversarial known-answer cases of the oracle tests also run on the GPU.
int tiab_dbg_watershed_device(tiab_handle_t h, const double* dist_host, const int32_t* markers_host, const uint8_t* mask_host, int H, int W,
                              int impl, int32_t* out_host) {
  if (!h || !dist_host || !markers_host || !mask_host || !out_host || H <= 0 || W <= 0 || (impl != 3 && impl != 4)) return TIAB_E_INVALID;
  TIAB_ON_DEVICE(h);
  const int P = H * W;
  std::vector<int> lab((size_t)P), size((size_t)P, 0), bbox((size_t)P * 4, 0), hoff((size_t)P, 0), mk((size_t)P);
  std::vector<uint8_t> blb((size_t)P);
  int x0 = W, y0 = H, x1 = -1, y1 = -1, npx = 0, r = -1;
  for (int p = 0; p < P; ++p) {
    blb[p] = mask_host[p] ? 1 : 0;
    mk[p] = mask_host[p] ? markers_host[p] : 0;
    if (mask_host[p]) {
      if (r < 0) r = p;
      ++npx;
      x0 = std::min(x0, p % W), x1 = std::max(x1, p % W), y0 = std::min(y0, p / W), y1 = std::max(y1, p / W);
    }
  }
  for (int p = 0; p < P; ++p) out_host[p] = 0;
  if (r < 0) return TIAB_OK;
  for (int p = 0; p < P; ++p) lab[p] = mask_host[p] ? r : -1;
  size[r] = npx;
  bbox[4 * r] = x0, bbox[4 * r + 1] = y0, bbox[4 * r + 2] = x1, bbox[4 * r + 3] = y1;
  const int wl[2] = {0, r}, wc[2] = {1, 0};
  int *d_lab, *d_size, *d_bbox, *d_hoff, *d_mk, *d_wl, *d_wc, *d_out;
  uint8_t* d_blb;
  double* d_dist;
  HeapItem* d_heap;
  TIAB_CUDA_OK(h, cudaMalloc(&d_lab, P * 4));
  TIAB_CUDA_OK(h, cudaMalloc(&d_size, P * 4));
  TIAB_CUDA_OK(h, cudaMalloc(&d_bbox, (size_t)P * 16));
  TIAB_CUDA_OK(h, cudaMalloc(&d_hoff, P * 4));
  TIAB_CUDA_OK(h, cudaMalloc(&d_mk, P * 4));
  TIAB_CUDA_OK(h, cudaMalloc(&d_wl, 8));
  TIAB_CUDA_OK(h, cudaMalloc(&d_wc, 8));
  TIAB_CUDA_OK(h, cudaMalloc(&d_out, P * 4));
  TIAB_CUDA_OK(h, cudaMalloc(&d_blb, P));
  TIAB_CUDA_OK(h, cudaMalloc(&d_dist, (size_t)P * 8));
  TIAB_CUDA_OK(h, cudaMalloc(&d_heap, (size_t)P * sizeof(HeapItem)));
  TIAB_CUDA_OK(h, cudaMemcpy(d_lab, lab.data(), P * 4, cudaMemcpyHostToDevice));
  TIAB_CUDA_OK(h, cudaMemcpy(d_size, size.data(), P * 4, cudaMemcpyHostToDevice));
  TIAB_CUDA_OK(h, cudaMemcpy(d_bbox, bbox.data(), (size_t)P * 16, cudaMemcpyHostToDevice));
  TIAB_CUDA_OK(h, cudaMemcpy(d_hoff, hoff.data(), P * 4, cudaMemcpyHostToDevice));
  TIAB_CUDA_OK(h, cudaMemcpy(d_mk, mk.data(), P * 4, cudaMemcpyHostToDevice));
  TIAB_CUDA_OK(h, cudaMemcpy(d_wl, wl, 8, cudaMemcpyHostToDevice));
  TIAB_CUDA_OK(h, cudaMemcpy(d_wc, wc, 8, cudaMemcpyHostToDevice));
  TIAB_CUDA_OK(h, cudaMemcpy(d_blb, blb.data(), P, cudaMemcpyHostToDevice));
  TIAB_CUDA_OK(h, cudaMemcpy(d_dist, dist_host, (size_t)P * 8, cudaMemcpyHostToDevice));
  TIAB_CUDA_OK(h, cudaMemset(d_out, 0, P * 4));
  const int smem_bytes = 200 * 1024;
  if (impl == 4) {
    TIAB_CUDA_OK(h, cudaFuncSetAttribute(pp_watershed_v4, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    pp_watershed_v4<<<1, 32, smem_bytes>>>(d_wl, d_wc, d_lab, d_size, d_bbox, d_hoff, d_blb, d_mk, d_dist, H, W, d_heap, d_out, smem_bytes);
  } else {
    TIAB_CUDA_OK(h, cudaFuncSetAttribute(pp_watershed_v3, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    pp_watershed_v3<<<1, 32, smem_bytes>>>(d_wl, d_wc, d_lab, d_size, d_bbox, d_hoff, d_blb, d_mk, d_dist, H, W, d_heap, d_out, smem_bytes);
  }
  TIAB_CUDA_OK(h, cudaGetLastError());
  TIAB_CUDA_OK(h, cudaMemcpy(out_host, d_out, P * 4, cudaMemcpyDeviceToHost));
  cudaFree(d_lab), cudaFree(d_size), cudaFree(d_bbox), cudaFree(d_hoff), cudaFree(d_mk), cudaFree(d_wl), cudaFree(d_wc), cudaFree(d_out);
  cudaFree(d_blb), cudaFree(d_dist), cudaFree(d_heap);
  return TIAB_OK;
}

int tiab_dbg_trace_contour(const int32_t* map_host, int H, int W, int32_t id, int* out_xy, int cap) {
  if (!map_host || H <= 0 || W <= 0 || cap < 0 || (cap && !out_xy)) return TIAB_E_INVALID;
  const LabelView v{map_host, H, W, id};
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x)
      if (v.fg(x, y)) return trace_outer_contour(v, x, y, out_xy, cap);
  return 0;
}

int tiab_dbg_sobel_kernels_k(int ksize, double* deriv, double* smooth) {
  if (ksize < 3 || ksize > 21 || !(ksize & 1) || !deriv || !smooth) return TIAB_E_INVALID;
  sobel_kernels(deriv, smooth, ksize);
  return TIAB_OK;
}

// copy an intermediate of the LAST tiab_hovernet_postproc call to the host (tests only):
// 0 blb u8, 1 h_dir f32, 2 v_dir f32, 3 sobel_h f64, 4 sobel_v f64, 5 dist f64, 6 marker i32, 7 opened seeds u8
int tiab_dbg_postproc_fetch(tiab_handle_t h, int which, void* host_out, void* stream) {
  if (!h || which < 0 || which > 7 || !host_out || !h->pp_work || !h->pp_dbg_bytes[which]) return TIAB_E_INVALID;
  TIAB_CUDA_OK(h, cudaMemcpyAsync(host_out, (uint8_t*)h->pp_work + h->pp_dbg_off[which], h->pp_dbg_bytes[which], cudaMemcpyDeviceToHost,
                                  (cudaStream_t)stream));
  TIAB_CUDA_OK(h, cudaStreamSynchronize((cudaStream_t)stream));
  return TIAB_OK;
}

int tiab_hovernet_postproc(tiab_handle_t h, const float* np_map, const float* hv_map, const float* tp_map, int N, int H, int W,
                           int32_t* inst_map, int max_inst, int* inst_count, int32_t* inst_id, int32_t* inst_box,
                           double* inst_centroid, int32_t* inst_type, double* inst_prob, void* stream) {
  return tiab_hovernet_postproc_scaled(h, np_map, hv_map, tp_map, N, H, W, 1.0, inst_map, max_inst, inst_count, inst_id, inst_box,
                                       inst_centroid, inst_type, inst_prob, stream);
}

int tiab_hovernet_postproc_scaled(tiab_handle_t h, const float* np_map, const float* hv_map, const float* tp_map, int N, int H, int W,
                                  double scale_factor, int32_t* inst_map, int max_inst, int* inst_count, int32_t* inst_id,
                                  int32_t* inst_box, double* inst_centroid, int32_t* inst_type, double* inst_prob, void* stream) {
  if (!h) return TIAB_E_INVALID;
  if (!np_map || !hv_map || !inst_map || N <= 0 || H < 11 || W < 11 || (long long)H * W > (1 << 24) || max_inst <= 0 || !(scale_factor > 0)) {
    h->err = "tiab_hovernet_postproc: invalid argument";
    return TIAB_E_INVALID;
  }
  // hovernet.py:564-565: ksize = int(20 * scale_factor + 1), obj_size = ceil(10 * scale_factor^2)
  const int ksize = (int)(20.0 * scale_factor + 1.0);
  const int marker_max_size = (int)ceil(10.0 * scale_factor * scale_factor) - 1;
  if (ksize != 21 && ksize != 11) {
    h->err = "tiab_hovernet_postproc: scale_factor must give a Sobel size of 21 (1.0: HoVer-Net) or 11 (0.5: HoVerNet+)";
    return TIAB_E_UNSUPPORTED;
  }
  const bool want_table = inst_count != nullptr;
  if (want_table && (!inst_id || !inst_box || !inst_centroid || !inst_type || !inst_prob)) {
    h->err = "tiab_hovernet_postproc: incomplete instance table outputs";
    return TIAB_E_INVALID;
  }
  TIAB_ON_DEVICE(h);
  cudaStream_t s = (cudaStream_t)stream;
  if (!h->pp_const_set) {
    double d[21], sm[21];
    sobel_kernels(d, sm, 21);
    TIAB_CUDA_OK(h, cudaMemcpyToSymbol(c_sobel_d, d, sizeof(d)));
    TIAB_CUDA_OK(h, cudaMemcpyToSymbol(c_sobel_s, sm, sizeof(sm)));
    sobel_kernels(d, sm, 11);
    TIAB_CUDA_OK(h, cudaMemcpyToSymbol(c_sobel_d11, d, 11 * sizeof(double)));
    TIAB_CUDA_OK(h, cudaMemcpyToSymbol(c_sobel_s11, sm, 11 * sizeof(double)));
    h->pp_const_set = 1;
  }
  const int P = H * W;
  const long long T = (long long)N * P;
  const long long TI = (long long)N * max_inst;
  // carve the workspace
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off = (off + bytes + 255) & ~(size_t)255;
    return o;
  };
  const size_t o_lab = take(T * 4), o_size = take(T * 4), o_cid = take(T * 4), o_roots = take(T * 4), o_ncomp = take((size_t)N * 4);
  const size_t o_bbox = take(T * 16), o_blb = take(T), o_fg = take(T), o_mk = take(T), o_tmp = take(T), o_border = take(T);
  const size_t o_hn = take(T * 4), o_vn = take(T * 4), o_rowh = take(T * 8), o_rowv = take(T * 8), o_sobh = take(T * 8), o_sobv = take(T * 8);
  const size_t o_dist0 = take(T * 8), o_dist = take(T * 8), o_mm = take((size_t)N * 16), o_mm64 = take((size_t)N * 32);
  const size_t o_marker = take(T * 4), o_heap = take(T * sizeof(HeapItem)), o_hoff = take(T * 4);
  const size_t o_roots2 = take(T * 4), o_ncomp2 = take((size_t)N * 4), o_lab2 = take(T * 4), o_size2 = take(T * 4);
  const size_t o_wlist = take(T * 4), o_wcount = take(64);   // flood v3: (image, root) pairs of the live components + {count, next}
  const size_t o_iarea = take(TI * 4), o_ibox = take(TI * 16), o_isx = take(TI * 8), o_isy = take(TI * 8), o_itype = take(TI * 4 * PP_MAX_TYPES);
  int rc = pp_ensure(h, off);
  if (rc) return rc;
  uint8_t* B = (uint8_t*)h->pp_work;
  PPWork w;
  w.lab = (int*)(B + o_lab), w.size = (int*)(B + o_size), w.cid = (int*)(B + o_cid), w.roots = (int*)(B + o_roots), w.ncomp = (int*)(B + o_ncomp);
  w.bbox = (int*)(B + o_bbox), w.blb = B + o_blb, w.fg = B + o_fg, w.mk = B + o_mk, w.tmp = B + o_tmp, w.border = B + o_border;
  w.hn = (float*)(B + o_hn), w.vn = (float*)(B + o_vn), w.rowh = (double*)(B + o_rowh), w.rowv = (double*)(B + o_rowv);
  w.sobh = (double*)(B + o_sobh), w.sobv = (double*)(B + o_sobv), w.dist0 = (double*)(B + o_dist0), w.dist = (double*)(B + o_dist);
  w.mm = (float*)(B + o_mm), w.mm64 = (double*)(B + o_mm64), w.marker = (int*)(B + o_marker), w.heap = (HeapItem*)(B + o_heap);
  w.heap_off = (int*)(B + o_hoff);
  int* roots2 = (int*)(B + o_roots2);
  int* ncomp2 = (int*)(B + o_ncomp2);
  int* lab2 = (int*)(B + o_lab2);
  int* size2 = (int*)(B + o_size2);
  w.iarea = (int*)(B + o_iarea), w.ibox = (int*)(B + o_ibox), w.isx = (unsigned long long*)(B + o_isx), w.isy = (unsigned long long*)(B + o_isy);
  w.itype = (int*)(B + o_itype);

  {
    const size_t offs[8] = {o_blb, o_hn, o_vn, o_sobh, o_sobv, o_dist, o_marker, o_mk};
    const size_t byt[8] = {(size_t)T, (size_t)T * 4, (size_t)T * 4, (size_t)T * 8, (size_t)T * 8, (size_t)T * 8, (size_t)T * 4, (size_t)T};
    for (int i = 0; i < 8; ++i) h->pp_dbg_off[i] = offs[i], h->pp_dbg_bytes[i] = byt[i];
  }
  h->last_launches = 0;
  const int g = (int)std::min<long long>((T + PP_THREADS - 1) / PP_THREADS, (long long)h->sm_count * 32);
  const dim3 g2(std::max(1, std::min(64, (P + PP_THREADS - 1) / PP_THREADS)), N);
  const int gN = (N + PP_THREADS - 1) / PP_THREADS;

  // 1. foreground: threshold -> CCL -> keep components > 9 px  (hovernet.py:541-545)
  PP_LAUNCH(h, pp_threshold, g, PP_THREADS, np_map, T, w.fg);
  PP_LAUNCH(h, pp_ccl_init, g, PP_THREADS, w.fg, P, W, T, w.lab, w.size, w.bbox, w.border);
  PP_LAUNCH(h, pp_ccl_merge, g, PP_THREADS, w.fg, H, W, T, w.lab);
  PP_LAUNCH(h, pp_ccl_flatten, g, PP_THREADS, H, W, T, w.lab, w.size, w.bbox, w.border);
  PP_LAUNCH(h, pp_rank_roots, N, 1024, w.lab, P, w.cid, w.roots, w.ncomp);
  PP_LAUNCH(h, pp_keep_large, g, PP_THREADS, w.lab, w.size, w.cid, P, T, 9, w.blb, (int*)nullptr);
  // 2. normalised gradients (hovernet.py:547-590)
  PP_LAUNCH(h, pp_mm_init, gN, PP_THREADS, N, w.mm, w.mm64);
  PP_LAUNCH(h, pp_minmax_hv, g2, PP_THREADS, hv_map, P, w.mm);
  PP_LAUNCH(h, pp_normalize_hv, g, PP_THREADS, hv_map, P, T, w.mm, w.hn, w.vn);
  if (ksize == 21) {
    PP_LAUNCH(h, pp_sobel_rows<21>, g, PP_THREADS, w.hn, w.vn, H, W, T, w.rowh, w.rowv);
    PP_LAUNCH(h, pp_sobel_cols<21>, g, PP_THREADS, w.rowh, w.rowv, H, W, T, w.sobh, w.sobv);
  } else {
    PP_LAUNCH(h, pp_sobel_rows<11>, g, PP_THREADS, w.hn, w.vn, H, W, T, w.rowh, w.rowv);
    PP_LAUNCH(h, pp_sobel_cols<11>, g, PP_THREADS, w.rowh, w.rowv, H, W, T, w.sobh, w.sobv);
  }
  PP_LAUNCH(h, pp_minmax_sobel, g2, PP_THREADS, w.sobh, w.sobv, P, w.mm64);
  // 3. energy, distance, seeds (hovernet.py:592-603)
  PP_LAUNCH(h, pp_energy, g, PP_THREADS, w.sobh, w.sobv, w.mm64, w.blb, P, T, w.dist0, w.mk);
  PP_LAUNCH(h, pp_blur_rows, g, PP_THREADS, w.dist0, W, T, w.rowh);
  PP_LAUNCH(h, pp_blur_cols, g, PP_THREADS, w.rowh, H, W, T, w.dist);
  // 4. fill holes (hovernet.py:604): CCL of the complement, holes = components that do not touch the border
  PP_LAUNCH(h, pp_invert, g, PP_THREADS, w.mk, T, w.fg);
  PP_LAUNCH(h, pp_ccl_init, g, PP_THREADS, w.fg, P, W, T, lab2, size2, (int*)nullptr, w.border);
  PP_LAUNCH(h, pp_ccl_merge, g, PP_THREADS, w.fg, H, W, T, lab2);
  PP_LAUNCH(h, pp_ccl_flatten, g, PP_THREADS, H, W, T, lab2, size2, (int*)nullptr, w.border);
  PP_LAUNCH(h, pp_fill_holes, g, PP_THREADS, w.mk, lab2, w.border, P, T, w.tmp);
  // 5. opening with the 5x5 ellipse (hovernet.py:605-606)
  PP_LAUNCH(h, pp_morph, g, PP_THREADS, w.tmp, H, W, T, 0, w.fg);
  PP_LAUNCH(h, pp_morph, g, PP_THREADS, w.fg, H, W, T, 1, w.mk);
  // 6. label the seeds, drop <= obj_size - 1 px (9 at scale 1), keep ids (hovernet.py:607-614)
  PP_LAUNCH(h, pp_ccl_init, g, PP_THREADS, w.mk, P, W, T, lab2, size2, (int*)nullptr, w.border);
  PP_LAUNCH(h, pp_ccl_merge, g, PP_THREADS, w.mk, H, W, T, lab2);
  PP_LAUNCH(h, pp_ccl_flatten, g, PP_THREADS, H, W, T, lab2, size2, (int*)nullptr, w.border);
  PP_LAUNCH(h, pp_rank_roots, N, 1024, lab2, P, w.cid, roots2, ncomp2);
  PP_LAUNCH(h, pp_keep_large, g, PP_THREADS, lab2, size2, w.cid, P, T, marker_max_size, (uint8_t*)nullptr, w.marker);
  // 7. watershed (hovernet.py:616)
  TIAB_CUDA_OK(h, cudaMemsetAsync(inst_map, 0, (size_t)T * 4, s));
  PP_LAUNCH(h, pp_heap_offsets, N, PP_THREADS, w.roots, w.ncomp, w.size, P, N, w.heap_off);
  {
    // one thread per possible component rank (ranks >= ncomp[n] and removed small components exit immediately)
    const int mc = P;
    const long long nthreads = (long long)N * mc;
    if (h->pp_watershed_impl == 3 || h->pp_watershed_impl == 4) {
      int* wlist = (int*)(B + o_wlist);
      int* wcount = (int*)(B + o_wcount);
      TIAB_CUDA_OK(h, cudaMemsetAsync(wcount, 0, 2 * sizeof(int), s));
      {
        const int cls[4] = {1 << 30, 1024, 128, 0};   // size classes, largest first (longest-processing-time-first scheduling)
        for (int k = 0; k < 3; ++k)
          PP_LAUNCH(h, pp_flood_worklist, (unsigned)((nthreads + 255) / 256), 256, w.roots, w.ncomp, w.size, P, N, mc, cls[k + 1], cls[k],
                    wlist, wcount);
      }
      const int smem_bytes = 200 * 1024;
      if (h->pp_watershed_impl == 4) {
        TIAB_CUDA_OK(h, cudaFuncSetAttribute(pp_watershed_v4, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));   // per device
        pp_watershed_v4<<<h->sm_count, 32, smem_bytes, s>>>(wlist, wcount, w.lab, w.size, w.bbox, w.heap_off, w.blb, w.marker, w.dist, H,
                                                            W, w.heap, inst_map, smem_bytes);
      } else {
        TIAB_CUDA_OK(h, cudaFuncSetAttribute(pp_watershed_v3, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));   // per device
        pp_watershed_v3<<<h->sm_count, 32, smem_bytes, s>>>(wlist, wcount, w.lab, w.size, w.bbox, w.heap_off, w.blb, w.marker, w.dist, H,
                                                            W, w.heap, inst_map, smem_bytes);
      }
      {
        cudaError_t _e = cudaGetLastError();
        if (_e != cudaSuccess) {
          h->err = std::string("pp_watershed_v3/v4 launch: ") + cudaGetErrorString(_e);
          return TIAB_E_CUDA;
        }
        h->last_launches++;
      }
    } else if (h->pp_watershed_impl == 2) {
      // only ranks below the largest component count can be live: bound the grid by min(P, a generous per-image cap) warps per image
      PP_LAUNCH(h, pp_watershed_v2, (unsigned)((nthreads * 32 + 127) / 128), 128, w.roots, w.ncomp, w.lab, w.size, w.bbox, w.heap_off, w.blb,
                w.marker, w.dist, H, W, N, mc, w.heap, inst_map);
    } else {
      PP_LAUNCH(h, pp_watershed, (unsigned)((nthreads + 63) / 64), 64, w.roots, w.ncomp, w.lab, w.size, w.bbox, w.heap_off, w.blb, w.marker,
                w.dist, H, W, N, mc, w.heap, inst_map);
    }
  }
  // 8. instance table (hovernet.py:670-748, without contours)
  if (want_table) {
    const int gi = (int)std::min<long long>((TI + PP_THREADS - 1) / PP_THREADS, (long long)h->sm_count * 32);
    PP_LAUNCH(h, pp_inst_init, gi, PP_THREADS, TI, w.iarea, w.ibox, w.isx, w.isy, w.itype);
    PP_LAUNCH(h, pp_inst_accum, g, PP_THREADS, inst_map, tp_map, H, W, T, max_inst, w.iarea, w.ibox, w.isx, w.isy, w.itype);
    PP_LAUNCH(h, pp_inst_finalize, N, PP_THREADS, N, max_inst, tp_map ? 1 : 0, w.iarea, w.ibox, w.isx, w.isy, w.itype, inst_count, inst_id,
              inst_box, inst_centroid, inst_type, inst_prob);
  }
  return TIAB_OK;
}

int tiab_hovernet_contours(tiab_handle_t h, const int32_t* inst_map, int N, int H, int W, int max_inst, const int* inst_count,
                           const int32_t* inst_id, const int32_t* inst_box, int capacity, int32_t* cont_off, int32_t* cont_len,
                           int32_t* cont_total, int32_t* cont_xy, void* stream) {
  if (!h) return TIAB_E_INVALID;
  if (!inst_map || N <= 0 || H <= 0 || W <= 0 || max_inst <= 0 || !inst_count || !inst_id || !inst_box || capacity <= 0 || !cont_off ||
      !cont_len || !cont_total || !cont_xy) {
    h->err = "tiab_hovernet_contours: invalid argument";
    return TIAB_E_INVALID;
  }
  TIAB_ON_DEVICE(h);
  cudaStream_t s = (cudaStream_t)stream;
  TIAB_CUDA_OK(h, cudaMemsetAsync(cont_total, 0, (size_t)N * sizeof(int32_t), s));
  const long long threads = (long long)N * max_inst;
  h->last_launches = 0;
  PP_LAUNCH(h, pp_contours, (unsigned)((threads + 127) / 128), 128, inst_map, N, H, W, max_inst, inst_count, inst_id, inst_box, capacity,
            cont_off, cont_len, cont_total, cont_xy);
  return TIAB_OK;
}

}  // extern "C"
