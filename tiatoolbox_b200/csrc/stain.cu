// stain.cu — B200 (sm_100a) stain-normalisation hot path.
//
// Two device paths, same arithmetic:
//   * FUSED   one persistent CTA per SM; a whole uint8 patch (<= 192 KB, e.g. 256x256x3) is staged
//             ONCE into shared memory with bulk-async (TMA 1-D) copies, every reduction of the
//             Macenko / Vahadane / Ruifrok pipeline (u8 histogram -> tissue mask + OD moments ->
//             3x3 eigen -> angular percentiles -> concentrations -> 99th percentiles -> recompose)
//             is recomputed from the SMEM-resident bytes, and the result leaves through one
//             bulk-async store.  HBM traffic = read once + write once = 2*H*W*3 bytes.
//   * GENERAL the same passes as separate kernels over global memory for images that do not fit
//             in shared memory (fit() on a 1000x1000 target, arbitrary H x W).
//
// Reference semantics: tools/stainnorm.py:49-113, tools/stainextract.py:177-227,281-322,
// utils/misc.py:261-290,405-444, utils/transforms.py:209-231 (see stain_math.cuh for the scalar
// steps and DESIGN.md for the pass structure and the order-statistic scheme).
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <string>
#include <thread>
#include <vector>

#include "../../include/tiab200.h"
#include "ctx.h"
#include "stain_math.cuh"

namespace tiab {

constexpr int NB = 1024;                 // bins per histogram level (two levels => range / 2^20)
constexpr int FUSED_THREADS = 1024;
constexpr int PATCH_BYTES_MAX = 196608;  // 256*256*3
constexpr int NCHUNK = 8;                // bulk-copy chunks per patch
constexpr float OD_MAX = 5.5412636f;     // -ln(1/255)
constexpr int HIST_WORDS = 4 * NB;       // shared histogram region (4 fine hists / 2 coarse / 8 u8 replicas)

enum { M_MACENKO = 0, M_VAHADANE = 1, M_RUIFROK = 2, M_CUSTOM = 3 };

struct StainTables {  // constant for the process, uploaded once per handle
  float od[256];
  unsigned tabg[256];
};

struct StainConst {  // per-call constants, passed by value
  int method;
  int H, W, P;
  int ysum_limit;  // tissue <=> ylut_r + ylut_g + ylut_b < ysum_limit
  int do_apply;    // 0: fit (stop after the concentration percentiles)
  double q_lo, q_hi;          // contrast_enhancer percentiles as fractions
  double q_phi_lo, q_phi_hi;  // Macenko angular percentiles as fractions
  double q_conc;              // 0.99
  double reg;                 // Vahadane regulariser
  double M_fixed[6];
  double M_target[6];
  double maxC_target[2];
};

struct StainState {  // small per-image state (shared memory in the fused kernel)
  int status;
  unsigned n_tissue;
  double p_lo, p_hi;
  double acc[12];  // block-reduced sums of the current pass
  double mom[10];
  double Ed[6];
  float E[6];
  double M_src[6];
  float A[6];
  float c_lo[2], c_scale[2];
  int tgt_bin[4];
  unsigned tgt_rank[4];
  double tgt_gamma[2];
  double tgt_val[4];
  double maxC_src[2];
  float W2[9];
  float vg[4][3], vbeta[4], vinv[4], vU[6];  // float copies of the Vahadane replay program
  int vsteps;
  VahState vah;
};

struct StainWork {  // per-image scratch of the GENERAL path (global memory)
  unsigned hist[HIST_WORDS];
  uint2 lut[768];
};

// ------------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned warp_sum_u(unsigned v) {
#pragma unroll
  for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}
__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}
__device__ __forceinline__ unsigned warp_incl_scan_u(unsigned v) {
  int lane = threadIdx.x & 31;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    unsigned n = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= d) v += n;
  }
  return v;
}

// Block-wide (1024 threads): locate the bin holding 0-based rank `rank` in an NB-bin histogram.
// scratch: >= 34 words of shared memory.  All threads get the result.
__device__ void block_find_rank(const unsigned* hist, unsigned rank, unsigned* scratch, int* bin, unsigned* within) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  unsigned s = warp_sum_u(hist[w * 32 + lane]);
  if (lane == 0) scratch[w] = s;
  __syncthreads();
  if (w == 0) {
    unsigned p = scratch[lane];
    unsigned incl = warp_incl_scan_u(p), excl = incl - p;
    unsigned m = __ballot_sync(0xffffffffu, rank >= excl && rank < incl);
    int chunk = m ? (__ffs(m) - 1) : 31;
    unsigned base = __shfl_sync(0xffffffffu, excl, chunk);
    unsigned v2 = hist[chunk * 32 + lane];
    unsigned incl2 = warp_incl_scan_u(v2), excl2 = incl2 - v2;
    unsigned r2 = rank >= base ? rank - base : 0u;
    unsigned m2 = __ballot_sync(0xffffffffu, r2 >= excl2 && r2 < incl2);
    int l2 = m2 ? (__ffs(m2) - 1) : 31;
    if (lane == l2) {
      scratch[32] = (unsigned)(chunk * 32 + l2);
      scratch[33] = m2 ? r2 - excl2 : 0u;
    }
  }
  __syncthreads();
  *bin = (int)scratch[32];
  *within = scratch[33];
  __syncthreads();
}

// K per-thread floats -> warp reduce -> double atomics into dst (shared or global).
template <int K>
__device__ __forceinline__ void block_accumulate(const float (&v)[K], double* dst) {
#pragma unroll
  for (int k = 0; k < K; ++k) {
    float s = warp_sum_f(v[k]);
    if ((threadIdx.x & 31) == 0) atomicAdd(dst + k, (double)s);
  }
}

template <class F>
__device__ __forceinline__ void quad_pixels(uint32_t w0, uint32_t w1, uint32_t w2, F&& f) {
  f(0, w0 & 255u, (w0 >> 8) & 255u, (w0 >> 16) & 255u);
  f(1, w0 >> 24, w1 & 255u, (w1 >> 8) & 255u);
  f(2, (w1 >> 16) & 255u, w1 >> 24, w2 & 255u);
  f(3, (w2 >> 8) & 255u, (w2 >> 16) & 255u, w2 >> 24);
}

// Pixel iteration shared by both paths.  `base` addresses P pixels (3 bytes each); when
// vec != 0 it is 4-byte aligned and P % 4 == 0 and 12-byte quads are read as 3 words.
template <class F>
__device__ __forceinline__ void for_pixels(const uint8_t* base, int vec, int px_begin, int px_end, int tid,
                                           int nthreads, F&& f) {
  if (vec) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(base);
    for (int q = (px_begin >> 2) + tid; q < (px_end >> 2); q += nthreads) {
      uint32_t w0 = w[3 * q], w1 = w[3 * q + 1], w2 = w[3 * q + 2];
      quad_pixels(w0, w1, w2, [&](int, unsigned r, unsigned g, unsigned b) { f(r, g, b); });
    }
  } else {
    for (int p = px_begin + tid; p < px_end; p += nthreads) f(base[3 * p], base[3 * p + 1], base[3 * p + 2]);
  }
}

// ------------------------------------------------------------------------------------------------
// per-pixel bodies
// ------------------------------------------------------------------------------------------------
struct PixelLut {
  const uint2* lut;  // [3][256] {od bits, Y contribution}
  const float* od;   // [256]
  int ysum_limit;
  __device__ __forceinline__ bool tissue(unsigned r, unsigned g, unsigned b, float& o0, float& o1, float& o2) const {
    uint2 a = lut[r], c = lut[256 + g], d = lut[512 + b];
    o0 = __uint_as_float(a.x);
    o1 = __uint_as_float(c.x);
    o2 = __uint_as_float(d.x);
    return (int)(a.y + c.y + d.y) < ysum_limit;
  }
};

__device__ __forceinline__ int coarse_bin(float u) {
  int b = (int)u;
  return b < 0 ? 0 : (b > NB - 1 ? NB - 1 : b);
}
__device__ __forceinline__ int fine_bin(float u, int cb) {
  int f = (int)((u - (float)cb) * (float)NB);
  return f < 0 ? 0 : (f > NB - 1 ? NB - 1 : f);
}
__device__ __forceinline__ float phi_key(const float* E, float o0, float o1, float o2) {
  float p0 = o0 * E[0] + o1 * E[2] + o2 * E[4];
  float p1 = o0 * E[1] + o1 * E[3] + o2 * E[5];
  return (pseudo_angle(p0, p1) + 2.0f) * (float)(NB / 4);
}
// replay the Vahadane atoms at a pixel; nfull = completed steps; if pending >= 0 also the raw
// (pre-normalisation) value of the pending step is returned in *pend.
__device__ __forceinline__ void vah_replay(const StainState* st, int nfull, bool pending, float o0, float o1,
                                           float o2, float& d0, float& d1, float& pend) {
  d0 = o0 * st->vU[0] + o1 * st->vU[2] + o2 * st->vU[4];
  d1 = o0 * st->vU[1] + o1 * st->vU[3] + o2 * st->vU[5];
  for (int s = 0; s < nfull; ++s) {
    float lin = o0 * st->vg[s][0] + o1 * st->vg[s][1] + o2 * st->vg[s][2];
    if ((s & 1) == 0)
      d0 = fmaxf(lin - st->vbeta[s] * d1, 0.f) * st->vinv[s];
    else
      d1 = fmaxf(lin - st->vbeta[s] * d0, 0.f) * st->vinv[s];
  }
  if (pending) {
    int s = nfull;
    float lin = o0 * st->vg[s][0] + o1 * st->vg[s][1] + o2 * st->vg[s][2];
    pend = fmaxf(lin - st->vbeta[s] * ((s & 1) == 0 ? d1 : d0), 0.f);
  }
}

// ------------------------------------------------------------------------------------------------
// block-level finalise steps (all FUSED_THREADS threads call; thread 0 does the float64 scalar work)
// ------------------------------------------------------------------------------------------------
// after the u8 histogram (merged into hist[0..255]): contrast range + per-channel LUTs
__device__ __noinline__ void fin_hist(const StainConst& C, const StainTables* tab, StainState* st, unsigned* hist,
                                      uint2* lut) {
  __syncthreads();
  if (threadIdx.x == 0) {
    double lo, hi;
    contrast_range(hist, 3u * (unsigned)C.P, C.q_lo, C.q_hi, &lo, &hi);
    st->p_lo = lo;
    st->p_hi = hi;
  }
  __syncthreads();
  const unsigned coef[3] = {871u, 2929u, 296u};
  for (int i = threadIdx.x; i < 768; i += blockDim.x) {
    int ch = i >> 8, v = i & 255;
    unsigned e = contrast_value(v, st->p_lo, st->p_hi);
    // "+2048 then >>12 <= ythr"  <=>  sum of the three contributions < ysum_limit
    lut[i] = make_uint2(__float_as_uint(tab->od[v]), coef[ch] * tab->tabg[e]);
  }
  __syncthreads();
}

__device__ __noinline__ void set_conc_ranges(StainState* st) {
  for (int j = 0; j < 2; ++j) {
    float lo = 0.f, hi = 0.f;
    for (int ch = 0; ch < 3; ++ch) {
      float a = st->A[ch * 2 + j];
      if (a < 0)
        lo += a * OD_MAX;
      else
        hi += a * OD_MAX;
    }
    float span = hi - lo;
    if (!(span > 1e-20f)) span = 1.f;
    lo -= 1e-4f * span;
    span *= 1.0002f;
    st->c_lo[j] = lo;
    st->c_scale[j] = (float)NB / span;
  }
}

// Given the source stain matrix: concentration operator A = pinv(M) and the coarse ranges.
__device__ __noinline__ void set_source_matrix(StainState* st, const double M[6]) {
  double A[6];
  for (int i = 0; i < 6; ++i) st->M_src[i] = M[i];
  if (!pinv23(M, A)) {
    st->status = TIAB_STAIN_DEGENERATE;
    return;
  }
  bool finite = true;
  for (int i = 0; i < 6; ++i) {
    st->A[i] = (float)A[i];
    finite = finite && isfinite(A[i]);
  }
  if (!finite) {
    st->status = TIAB_STAIN_DEGENERATE;
    return;
  }
  set_conc_ranges(st);
}

// after the mask + moments pass (acc = {n, s[3], ss[6]})
__device__ __noinline__ void fin_moments(const StainConst& C, StainState* st) {
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 0; i < 10; ++i) st->mom[i] = st->acc[i];
    st->n_tissue = (unsigned)(st->acc[0] + 0.5);
    if (st->n_tissue == 0)
      st->status = TIAB_STAIN_EMPTY_TISSUE_MASK;
    else if (st->n_tissue < 2)
      st->status = TIAB_STAIN_DEGENERATE;
    else if (C.method == M_MACENKO) {
      macenko_basis(st->mom, st->Ed);
      for (int i = 0; i < 6; ++i) st->E[i] = (float)st->Ed[i];
    } else if (C.method == M_VAHADANE) {
      vah_init(&st->vah, st->mom, C.reg);
      for (int i = 0; i < 6; ++i) st->vU[i] = (float)st->vah.U[i];
      st->vsteps = 0;
    }
  }
  __syncthreads();
}

// rank targets for a percentile q over n samples: targets t0 (rank k) and t0+1 (rank k+1)
__device__ __forceinline__ void set_rank_targets(StainState* st, int t0, int gi, unsigned n, double q) {
  double vidx = dmul((double)(n - 1), q);
  double prev = floor(vidx);
  st->tgt_gamma[gi] = dsub(vidx, prev);
  unsigned k = (unsigned)prev;
  st->tgt_rank[t0] = k;
  st->tgt_rank[t0 + 1] = k + 1 < n ? k + 1 : n - 1;
}

// coarse histograms -> the coarse bin + rank-within-bin of each of the 4 targets.
// hist_of_target[t] selects which coarse histogram (0 or 1) target t lives in.
__device__ __noinline__ void fin_coarse(StainState* st, const unsigned* hist, int h01, int h23, unsigned* scratch) {
  for (int t = 0; t < 4; ++t) {
    int bin;
    unsigned within;
    block_find_rank(hist + (t < 2 ? h01 : h23) * NB, st->tgt_rank[t], scratch, &bin, &within);
    if (threadIdx.x == 0) {
      st->tgt_bin[t] = bin;
      st->tgt_rank[t] = within;
    }
  }
  __syncthreads();
}
// fine histograms (one per target) -> tgt_val[t] = position in coarse-bin units
__device__ __noinline__ void fin_fine(StainState* st, const unsigned* hist, unsigned* scratch) {
  for (int t = 0; t < 4; ++t) {
    int bin;
    unsigned within;
    block_find_rank(hist + t * NB, st->tgt_rank[t], scratch, &bin, &within);
    if (threadIdx.x == 0) st->tgt_val[t] = (double)st->tgt_bin[t] + ((double)bin + 0.5) / (double)NB;
  }
  __syncthreads();
}

__device__ __noinline__ void fin_phi(const StainConst& C, StainState* st) {
  __syncthreads();
  if (threadIdx.x == 0) {
    double ang[4];
    for (int t = 0; t < 4; ++t) ang[t] = angle_of_pseudo(st->tgt_val[t] * (4.0 / NB) - 2.0);
    double phi_min = np_lerp(ang[0], ang[1], st->tgt_gamma[0]);
    double phi_max = np_lerp(ang[2], ang[3], st->tgt_gamma[1]);
    double M[6];
    macenko_from_angles(st->Ed, phi_min, phi_max, M);
    set_source_matrix(st, M);
  }
  __syncthreads();
}

__device__ __noinline__ void fin_conc(const StainConst& C, StainState* st) {
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int j = 0; j < 2; ++j) {
      double a = (double)st->c_lo[j] + st->tgt_val[2 * j] / (double)st->c_scale[j];
      double b = (double)st->c_lo[j] + st->tgt_val[2 * j + 1] / (double)st->c_scale[j];
      st->maxC_src[j] = np_lerp(a, b, st->tgt_gamma[j]);
    }
    // out_c = 255 * exp(-(C0*s0*Mt[0][c] + C1*s1*Mt[1][c])),  C_j = od . A[:, j]
    const double log2e = 1.4426950408889634;
    for (int ch = 0; ch < 3; ++ch)
      for (int c = 0; c < 3; ++c) {
        double w = 0;
        for (int j = 0; j < 2; ++j)
          w += (double)st->A[ch * 2 + j] * (C.maxC_target[j] / st->maxC_src[j]) * C.M_target[j * 3 + c];
        st->W2[ch * 3 + c] = (float)(-w * log2e);
      }
    bool ok = isfinite(st->maxC_src[0]) && isfinite(st->maxC_src[1]) && st->maxC_src[0] != 0.0 && st->maxC_src[1] != 0.0;
    if (!ok && st->status == TIAB_STAIN_OK) st->status = TIAB_STAIN_DEGENERATE;
  }
  __syncthreads();
}

// Vahadane: start the next dictionary step (or finish).  Returns (to all threads, via state):
//   vah.done == 1  -> stain matrix set.   Otherwise a pass is pending:
//   phase even (norm pass for step vsteps) / odd (G,C pass).
__device__ __noinline__ void vah_prepare_step(const StainConst& C, StainState* st) {
  __syncthreads();
  if (threadIdx.x == 0 && st->status == TIAB_STAIN_OK) {
    int k = st->vah.nsteps & 1;
    if (!vah_begin_step(&st->vah, k)) {
      st->status = TIAB_STAIN_UNUSED_ATOM;
    } else {
      int s = st->vah.nsteps;
      for (int ch = 0; ch < 3; ++ch) st->vg[s][ch] = (float)st->vah.g[s][ch];
      st->vbeta[s] = (float)st->vah.beta[s];
    }
    for (int i = 0; i < 12; ++i) st->acc[i] = 0.0;
  }
  __syncthreads();
}
__device__ __noinline__ void vah_finish_norm(StainState* st) {
  __syncthreads();
  if (threadIdx.x == 0 && st->status == TIAB_STAIN_OK) {
    int s = st->vah.nsteps;
    vah_end_step(&st->vah, st->acc[0]);
    st->vinv[s] = (float)st->vah.inv[s];
    st->vsteps = st->vah.nsteps;
    for (int i = 0; i < 12; ++i) st->acc[i] = 0.0;
  }
  __syncthreads();
}
__device__ __noinline__ void vah_finish_gc(const StainConst& C, StainState* st) {
  __syncthreads();
  if (threadIdx.x == 0 && st->status == TIAB_STAIN_OK) {
    vah_after_gc(&st->vah, st->acc, C.reg, st->mom[0]);
    if (st->vah.done) {
      double M[6];
      vah_matrix(&st->vah, M);
      bool fin = true;
      for (int i = 0; i < 6; ++i) fin = fin && isfinite(M[i]);
      if (!fin)
        st->status = TIAB_STAIN_DEGENERATE;
      else
        set_source_matrix(st, M);
    }
  }
  __syncthreads();
}

// Exact order statistics inside the selected coarse bins (fused path): the few hundred keys that fall into a target's
// coarse bin are appended to a shared-memory list during the second pass and ranked by counting.  `lists` holds up to
// 4 lists of LIST_CAP floats (the histogram region); target t uses list lst[t]; tgt_rank[t] is its rank inside the bin.
constexpr int LIST_CAP = NB;
__device__ __noinline__ void fin_exact_select(StainState* st, const float* lists, const unsigned* cnt, const int* lst) {
  __syncthreads();
  for (int l = 0; l < 4; ++l) {
    bool used = false;
    for (int t = 0; t < 4; ++t) used = used || (lst[t] == l);
    if (!used) continue;
    const unsigned n = cnt[l] < (unsigned)LIST_CAP ? cnt[l] : (unsigned)LIST_CAP;
    const float* L = lists + l * LIST_CAP;
    for (unsigned i = threadIdx.x; i < n; i += blockDim.x) {
      const float v = L[i];
      unsigned less = 0, eq = 0;
      for (unsigned j = 0; j < n; ++j) {
        const float w = L[j];
        less += (w < v) ? 1u : 0u;
        eq += (w == v) ? 1u : 0u;
      }
      for (int t = 0; t < 4; ++t)
        if (lst[t] == l) {
          const unsigned r = st->tgt_rank[t];
          if (less <= r && r < less + eq) st->tgt_val[t] = (double)v;
        }
    }
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// bulk-async (TMA 1-D) copy + mbarrier primitives
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// FUSED kernel (v2)
//   * optical density comes from MUFU.LG2 on the byte (no shared-memory LUT, no bank conflicts):
//     L = log2(v) ; t = log2(255) - L = OD / ln2.  All linear forms (projections, concentrations,
//     output exponents) are pre-scaled on the host side of each pass so the pixel loops consume L.
//   * the tissue mask is computed once (pass B) and kept in two registers per thread (the same
//     thread owns the same 16 quads = 64 pixels in every pass).
//   * per-pixel bodies are branch-free; histogram atomics are predicated.
//   r1a profile of v1 (profiles/r1a_stain_fused_v1.md): 67 % shared-pipe utilisation, 55 % of the
//   wavefronts bank-conflict replays of LUT look-ups -> this rewrite.
// ------------------------------------------------------------------------------------------------
struct FusedSmem {
  // offsets into dynamic shared memory
  static constexpr int PATCH = 0;
  static constexpr int HIST = PATCH_BYTES_MAX;
  static constexpr int LUT = HIST + HIST_WORDS * 4;   // uint2[768] (general-path layout, built by fin_hist)
  static constexpr int YL = LUT + 768 * 8;            // unsigned[768]: luminance contributions only
  static constexpr int STATE = YL + 768 * 4;
  static constexpr int SCRATCH = STATE + 2560;
  static constexpr int MBAR = SCRATCH + 64 * 4;
  static constexpr int TOTAL = MBAR + NCHUNK * 8;
};
static_assert(sizeof(StainState) <= 2560, "StainState too large for its shared-memory slot");
static_assert(FusedSmem::TOTAL <= 232448, "shared memory budget exceeded");

__device__ __forceinline__ void zero_words(unsigned* p, int n) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) p[i] = 0u;
}

constexpr float LOG2_255 = 7.9943534368588578f;
constexpr float LN2_F = 0.69314718055994531f;

// single-instruction MUFU wrappers (the CUDA intrinsics add denormal-range fix-up code around the MUFU op)
__device__ __forceinline__ float mufu_lg2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float mufu_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float mufu_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// log2 of byte `sel & 3` of w: PRMT builds 0x4B0000vv (= 8388608 + v as a float), one FADD, one MUFU.LG2
__device__ __forceinline__ float lg_byte(uint32_t w, uint32_t sel) {
  return mufu_lg2(__uint_as_float(__byte_perm(w, 0x4B000000u, sel)) - 8388608.f);
}
__device__ __forceinline__ void quad_logs(uint32_t w0, uint32_t w1, uint32_t w2, float (&L)[12]) {
  L[0] = lg_byte(w0, 0x7650u), L[1] = lg_byte(w0, 0x7651u), L[2] = lg_byte(w0, 0x7652u), L[3] = lg_byte(w0, 0x7653u);
  L[4] = lg_byte(w1, 0x7650u), L[5] = lg_byte(w1, 0x7651u), L[6] = lg_byte(w1, 0x7652u), L[7] = lg_byte(w1, 0x7653u);
  L[8] = lg_byte(w2, 0x7650u), L[9] = lg_byte(w2, 0x7651u), L[10] = lg_byte(w2, 0x7652u), L[11] = lg_byte(w2, 0x7653u);
}
__device__ __forceinline__ unsigned byte_of(uint32_t w, int b) { return (w >> (8 * b)) & 255u; }

// pseudo-angle key in coarse-bin units, branch-free (see pseudo_angle() in stain_math.cuh)
__device__ __forceinline__ float phi_key_fast(float p0, float p1) {
  float s = fabsf(p0) + fabsf(p1);
  float r = p1 * mufu_rcp(s);
  float alt = copysignf(2.f, p1) - r;
  r = p0 < 0.f ? alt : r;
  return (r + 2.0f) * (float)(NB / 4);
}

struct LinForms {  // affine forms of L = log2(byte) used by the pixel loops
  float e[6], pb[2];   // projections: p_k = pb_k + sum_ch L_ch * e[ch*2+k]
  float a[6], cb[2];   // concentrations (already in coarse-bin units): u_j = cb_j + sum_ch L_ch * a[ch*2+j]
  float w[9], eb[3];   // output exponents (base 2): x_c = eb_c + sum_ch L_ch * w[ch*3+c]
};

__global__ void __launch_bounds__(FUSED_THREADS, 1)
stain_fused_kernel(const StainConst C, const StainTables* __restrict__ tab, const uint8_t* __restrict__ in,
                   uint8_t* __restrict__ out, int N, double* __restrict__ M_src_out, double* __restrict__ maxC_out,
                   int* __restrict__ status_out, float* __restrict__ out_f32) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint32_t* s_patch = reinterpret_cast<uint32_t*>(smem + FusedSmem::PATCH);
  unsigned* s_hist = reinterpret_cast<unsigned*>(smem + FusedSmem::HIST);
  uint2* s_lut = reinterpret_cast<uint2*>(smem + FusedSmem::LUT);
  unsigned* s_yl = reinterpret_cast<unsigned*>(smem + FusedSmem::YL);
  StainState* st = reinterpret_cast<StainState*>(smem + FusedSmem::STATE);
  unsigned* s_scratch = reinterpret_cast<unsigned*>(smem + FusedSmem::SCRATCH);
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + FusedSmem::MBAR);

  const int tid = threadIdx.x;
  const int P = C.P;
  const int nquads = P >> 2;
  const unsigned patch_bytes = (unsigned)P * 3u;
  const unsigned chunk_bytes = patch_bytes / NCHUNK;
  const int chunk_quads = nquads / NCHUNK;

  if (tid == 0) {
    for (int c = 0; c < NCHUNK; ++c) mbar_init(&s_bar[c], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  unsigned parity = 0;
  for (int n = blockIdx.x; n < N; n += gridDim.x, parity ^= 1u) {
    const uint8_t* src = in + (size_t)n * patch_bytes;
    uint8_t* dst = out + (size_t)n * patch_bytes;
    // ---- stage the patch: NCHUNK bulk copies, each signalling its own mbarrier ----
    if (tid == 0) {
      bulk_wait_read0();  // previous patch's stores have finished reading shared memory
      for (int c = 0; c < NCHUNK; ++c) {
        mbar_expect_tx(&s_bar[c], chunk_bytes);
        bulk_g2s(smem + (size_t)c * chunk_bytes, src + (size_t)c * chunk_bytes, chunk_bytes, &s_bar[c]);
      }
      st->status = TIAB_STAIN_OK;
      st->n_tissue = 0;
      for (int i = 0; i < 12; ++i) st->acc[i] = 0.0;
    }
    const bool need_mask = (C.method == M_MACENKO || C.method == M_VAHADANE);
    uint32_t mlo = 0u, mhi = 0u;  // tissue bits of this thread's pixels: quad iteration `it`, pixel j -> bit 4*it + j
    bool has_zero = true;

    if (need_mask) {
      // ---- pass A: exact histogram of all 3P uint8 values, overlapped with the copy ----
      zero_words(s_hist, 8 * 256);
      __syncthreads();
      unsigned* myhist = s_hist + ((tid >> 5) & 7) * 256;
      for (int c = 0; c < NCHUNK; ++c) {
        mbar_wait(&s_bar[c], parity);
        const int w_begin = c * chunk_quads * 3, w_end = w_begin + chunk_quads * 3;
        for (int w = w_begin + tid; w < w_end; w += FUSED_THREADS) {
          uint32_t v = s_patch[w];
          atomicAdd(&myhist[v & 255u], 1u);
          atomicAdd(&myhist[(v >> 8) & 255u], 1u);
          atomicAdd(&myhist[(v >> 16) & 255u], 1u);
          atomicAdd(&myhist[v >> 24], 1u);
        }
      }
      __syncthreads();
      if (tid < 256) {
        unsigned s = 0;
#pragma unroll
        for (int r = 0; r < 8; ++r) s += s_hist[r * 256 + tid];
        s_hist[tid] = s;  // replica-0 slot; only this thread ever reads bin `tid` of any replica
      }
      __syncthreads();
      has_zero = s_hist[0] != 0u;
      fin_hist(C, tab, st, s_hist, s_lut);
      for (int i = tid; i < 768; i += FUSED_THREADS) s_yl[i] = s_lut[i].y;
      __syncthreads();

      // ---- pass B: tissue mask (kept in registers) + moments of t = OD/ln2 over tissue pixels ----
      {
        float a[10];
#pragma unroll
        for (int i = 0; i < 10; ++i) a[i] = 0.f;
        const int ylimit = C.ysum_limit;
        for (int q = tid, it = 0; q < nquads; q += FUSED_THREADS, ++it) {
          const uint32_t w0 = s_patch[3 * q], w1 = s_patch[3 * q + 1], w2 = s_patch[3 * q + 2];
          unsigned vb[12];
#pragma unroll
          for (int b = 0; b < 4; ++b) vb[b] = byte_of(w0, b), vb[4 + b] = byte_of(w1, b), vb[8 + b] = byte_of(w2, b);
          unsigned nib = 0u;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int ysum = (int)(s_yl[vb[3 * j]] + s_yl[256 + vb[3 * j + 1]] + s_yl[512 + vb[3 * j + 2]]);
            const bool m = ysum < ylimit;
            nib |= (m ? 1u : 0u) << j;
            const float mf = m ? 1.f : 0.f;
            const float t0 = LOG2_255 - mufu_lg2(fmaxf(__uint_as_float(0x4B000000u | vb[3 * j]) - 8388608.f, 1.f));
            const float t1 = LOG2_255 - mufu_lg2(fmaxf(__uint_as_float(0x4B000000u | vb[3 * j + 1]) - 8388608.f, 1.f));
            const float t2 = LOG2_255 - mufu_lg2(fmaxf(__uint_as_float(0x4B000000u | vb[3 * j + 2]) - 8388608.f, 1.f));
            const float m0 = t0 * mf, m1 = t1 * mf, m2 = t2 * mf;
            a[0] += mf, a[1] += m0, a[2] += m1, a[3] += m2;
            a[4] = fmaf(m0, t0, a[4]), a[5] = fmaf(m0, t1, a[5]), a[6] = fmaf(m0, t2, a[6]);
            a[7] = fmaf(m1, t1, a[7]), a[8] = fmaf(m1, t2, a[8]), a[9] = fmaf(m2, t2, a[9]);
          }
          if (it < 8)
            mlo |= nib << (4 * it);
          else
            mhi |= nib << (4 * (it - 8));
        }
        block_accumulate<10>(a, st->acc);
      }
      __syncthreads();
      if (tid == 0) {  // t-moments -> OD moments (OD = ln2 * t)
        const double l = 0.6931471805599453, l2 = l * l;
        for (int i = 1; i < 4; ++i) st->acc[i] *= l;
        for (int i = 4; i < 10; ++i) st->acc[i] *= l2;
      }
      fin_moments(C, st);
    } else {
      for (int c = 0; c < NCHUNK; ++c) mbar_wait(&s_bar[c], parity);
      if (tid == 0) set_source_matrix(st, C.M_fixed);
      __syncthreads();
    }

    // zeros count as ones for the optical density (utils/transforms.py:229-230): fix them once in shared
    // memory (after the mask, which the reference computes on the unmodified image) so that the
    // remaining passes need no clamp.
    if (has_zero && st->status == TIAB_STAIN_OK) {
      for (int w = tid; w < nquads * 3; w += FUSED_THREADS) {
        const uint32_t v = s_patch[w];
        const uint32_t z = (v - 0x01010101u) & ~v & 0x80808080u;
        if (z) s_patch[w] = v | (z >> 7);
      }
      __syncthreads();
    }

    auto tissue_nib = [&](int it) -> unsigned { return ((it < 8 ? mlo : mhi) >> (4 * (it & 7))) & 15u; };

    if (need_mask && C.method == M_MACENKO && st->status == TIAB_STAIN_OK) {
      // ---- pass C: coarse histogram of the pseudo-angle over tissue pixels ----
      zero_words(s_hist, NB);
      if (tid == 0) {
        set_rank_targets(st, 0, 0, st->n_tissue, C.q_phi_lo);
        set_rank_targets(st, 2, 1, st->n_tissue, C.q_phi_hi);
      }
      __syncthreads();
      float e[6], pb[2];
      {
        float se0 = 0.f, se1 = 0.f;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
          e[ch * 2] = -LN2_F * st->E[ch * 2], e[ch * 2 + 1] = -LN2_F * st->E[ch * 2 + 1];
          se0 += st->E[ch * 2], se1 += st->E[ch * 2 + 1];
        }
        pb[0] = LOG2_255 * LN2_F * se0, pb[1] = LOG2_255 * LN2_F * se1;
      }
      for (int q = tid, it = 0; q < nquads; q += FUSED_THREADS, ++it) {
        float L[12];
        quad_logs(s_patch[3 * q], s_patch[3 * q + 1], s_patch[3 * q + 2], L);
        const unsigned nib = tissue_nib(it);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float p0 = fmaf(L[3 * j], e[0], fmaf(L[3 * j + 1], e[2], fmaf(L[3 * j + 2], e[4], pb[0])));
          const float p1 = fmaf(L[3 * j], e[1], fmaf(L[3 * j + 1], e[3], fmaf(L[3 * j + 2], e[5], pb[1])));
          const int cbn = coarse_bin(phi_key_fast(p0, p1));
          if ((nib >> j) & 1u) atomicAdd(&s_hist[cbn], 1u);
        }
      }
      __syncthreads();
      fin_coarse(st, s_hist, 0, 0, s_scratch);
      // ---- pass D: exact selection inside the selected coarse bins (candidate lists; histogram fallback) ----
      int tb[4], lst[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) tb[t] = st->tgt_bin[t];
      lst[0] = 0, lst[1] = tb[1] == tb[0] ? 0 : 1, lst[2] = 2, lst[3] = tb[3] == tb[2] ? 2 : 3;
      float* s_list = reinterpret_cast<float*>(s_hist);
      unsigned* s_cnt = s_scratch + 40;
      for (int attempt = 0; attempt < 2; ++attempt) {
        if (attempt == 1) zero_words(s_hist, 4 * NB);
        if (tid < 4) s_cnt[tid] = 0u;
        __syncthreads();
        for (int q = tid, it = 0; q < nquads; q += FUSED_THREADS, ++it) {
          float L[12];
          quad_logs(s_patch[3 * q], s_patch[3 * q + 1], s_patch[3 * q + 2], L);
          const unsigned nib = tissue_nib(it);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float p0 = fmaf(L[3 * j], e[0], fmaf(L[3 * j + 1], e[2], fmaf(L[3 * j + 2], e[4], pb[0])));
            const float p1 = fmaf(L[3 * j], e[1], fmaf(L[3 * j + 1], e[3], fmaf(L[3 * j + 2], e[5], pb[1])));
            const float u = phi_key_fast(p0, p1);
            const int cbn = coarse_bin(u);
            const bool m = (nib >> j) & 1u;
            if (!(m && (cbn == tb[0] || cbn == tb[1] || cbn == tb[2] || cbn == tb[3]))) continue;  // rare path below
#pragma unroll
            for (int t = 0; t < 4; ++t)
              if (cbn == tb[t] && lst[t] == t) {
                if (attempt == 0) {
                  const unsigned i = atomicAdd(&s_cnt[t], 1u);
                  if (i < (unsigned)LIST_CAP) s_list[t * LIST_CAP + i] = u;
                } else {
                  atomicAdd(&s_hist[t * NB + fine_bin(u, cbn)], 1u);
                }
              }
          }
        }
        __syncthreads();
        if (attempt == 0) {
          const bool overflow = s_cnt[0] > LIST_CAP || s_cnt[1] > LIST_CAP || s_cnt[2] > LIST_CAP || s_cnt[3] > LIST_CAP;
          if (!overflow) {
            fin_exact_select(st, s_list, s_cnt, lst);
            break;
          }
          __syncthreads();
        } else {
          // histogram fallback: duplicate lists share their histogram
          if (lst[1] != 1)
            for (int i = tid; i < NB; i += FUSED_THREADS) s_hist[1 * NB + i] = s_hist[0 * NB + i];
          if (lst[3] != 3)
            for (int i = tid; i < NB; i += FUSED_THREADS) s_hist[3 * NB + i] = s_hist[2 * NB + i];
          __syncthreads();
          fin_fine(st, s_hist, s_scratch);
        }
      }
      fin_phi(C, st);
    } else if (need_mask && C.method == M_VAHADANE && st->status == TIAB_STAIN_OK) {
      // ---- Vahadane: <= 2 dictionary updates, each = 2 norm passes + 1 (G, C) pass ----
      while (st->status == TIAB_STAIN_OK && !st->vah.done) {
        for (int half = 0; half < 2 && st->status == TIAB_STAIN_OK; ++half) {
          vah_prepare_step(C, st);
          if (st->status != TIAB_STAIN_OK) break;
          const int nfull = st->vsteps;
          float a[1] = {0.f};
          for (int q = tid, it = 0; q < nquads; q += FUSED_THREADS, ++it) {
            float L[12];
            quad_logs(s_patch[3 * q], s_patch[3 * q + 1], s_patch[3 * q + 2], L);
            const unsigned nib = tissue_nib(it);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float o0 = fmaf(L[3 * j], -LN2_F, LOG2_255 * LN2_F), o1 = fmaf(L[3 * j + 1], -LN2_F, LOG2_255 * LN2_F),
                          o2 = fmaf(L[3 * j + 2], -LN2_F, LOG2_255 * LN2_F);
              float d0, d1, pend;
              vah_replay(st, nfull, true, o0, o1, o2, d0, d1, pend);
              a[0] += ((nib >> j) & 1u) ? pend * pend : 0.f;
            }
          }
          block_accumulate<1>(a, st->acc);
          __syncthreads();
          vah_finish_norm(st);
        }
        if (st->status != TIAB_STAIN_OK) break;
        {
          const int nfull = st->vsteps;
          float a[9];
#pragma unroll
          for (int i = 0; i < 9; ++i) a[i] = 0.f;
          for (int q = tid, it = 0; q < nquads; q += FUSED_THREADS, ++it) {
            float L[12];
            quad_logs(s_patch[3 * q], s_patch[3 * q + 1], s_patch[3 * q + 2], L);
            const unsigned nib = tissue_nib(it);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float o0 = fmaf(L[3 * j], -LN2_F, LOG2_255 * LN2_F), o1 = fmaf(L[3 * j + 1], -LN2_F, LOG2_255 * LN2_F),
                          o2 = fmaf(L[3 * j + 2], -LN2_F, LOG2_255 * LN2_F);
              float d0, d1, pend;
              vah_replay(st, nfull, false, o0, o1, o2, d0, d1, pend);
              const float mf = ((nib >> j) & 1u) ? 1.f : 0.f;
              d0 *= mf, d1 *= mf;
              a[0] = fmaf(d0, d0, a[0]), a[1] = fmaf(d0, d1, a[1]), a[2] = fmaf(d1, d1, a[2]);
              a[3] = fmaf(d0, o0, a[3]), a[4] = fmaf(d0, o1, a[4]), a[5] = fmaf(d0, o2, a[5]);
              a[6] = fmaf(d1, o0, a[6]), a[7] = fmaf(d1, o1, a[7]), a[8] = fmaf(d1, o2, a[8]);
            }
          }
          block_accumulate<9>(a, st->acc);
          __syncthreads();
          vah_finish_gc(C, st);
        }
      }
    }

    if (st->status == TIAB_STAIN_OK) {
      // ---- pass E: coarse histograms of both stain concentrations over ALL pixels ----
      zero_words(s_hist, 2 * NB);
      if (tid == 0) {
        set_rank_targets(st, 0, 0, (unsigned)P, C.q_conc);
        set_rank_targets(st, 2, 1, (unsigned)P, C.q_conc);
      }
      __syncthreads();
      // u_j = (c_j - lo_j) * scale_j with c_j = sum_ch ln2*(K - L_ch) * A[ch][j]
      float a[6], cbias[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        float sa = 0.f;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
          a[ch * 2 + j] = -LN2_F * st->A[ch * 2 + j] * st->c_scale[j];
          sa += st->A[ch * 2 + j];
        }
        cbias[j] = (LOG2_255 * LN2_F * sa - st->c_lo[j]) * st->c_scale[j];
      }
      for (int q = tid; q < nquads; q += FUSED_THREADS) {
        float L[12];
        quad_logs(s_patch[3 * q], s_patch[3 * q + 1], s_patch[3 * q + 2], L);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float u0 = fmaf(L[3 * j], a[0], fmaf(L[3 * j + 1], a[2], fmaf(L[3 * j + 2], a[4], cbias[0])));
          const float u1 = fmaf(L[3 * j], a[1], fmaf(L[3 * j + 1], a[3], fmaf(L[3 * j + 2], a[5], cbias[1])));
          atomicAdd(&s_hist[coarse_bin(u0)], 1u);
          atomicAdd(&s_hist[NB + coarse_bin(u1)], 1u);
        }
      }
      __syncthreads();
      fin_coarse(st, s_hist, 0, 1, s_scratch);
      // ---- pass F: exact selection inside the selected coarse bins (candidate lists; histogram fallback) ----
      int tb[4], lst[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) tb[t] = st->tgt_bin[t];
      lst[0] = 0, lst[1] = tb[1] == tb[0] ? 0 : 1, lst[2] = 2, lst[3] = tb[3] == tb[2] ? 2 : 3;
      float* s_list = reinterpret_cast<float*>(s_hist);
      unsigned* s_cnt = s_scratch + 40;
      for (int attempt = 0; attempt < 2; ++attempt) {
        if (attempt == 1) zero_words(s_hist, 4 * NB);
        if (tid < 4) s_cnt[tid] = 0u;
        __syncthreads();
        for (int q = tid; q < nquads; q += FUSED_THREADS) {
          float L[12];
          quad_logs(s_patch[3 * q], s_patch[3 * q + 1], s_patch[3 * q + 2], L);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float u0 = fmaf(L[3 * j], a[0], fmaf(L[3 * j + 1], a[2], fmaf(L[3 * j + 2], a[4], cbias[0])));
            const float u1 = fmaf(L[3 * j], a[1], fmaf(L[3 * j + 1], a[3], fmaf(L[3 * j + 2], a[5], cbias[1])));
            const int cb0 = coarse_bin(u0), cb1 = coarse_bin(u1);
            if (!(cb0 == tb[0] || cb0 == tb[1] || cb1 == tb[2] || cb1 == tb[3])) continue;  // rare path below
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const float u = t < 2 ? u0 : u1;
              const int cbn = t < 2 ? cb0 : cb1;
              if (cbn == tb[t] && lst[t] == t) {
                if (attempt == 0) {
                  const unsigned i = atomicAdd(&s_cnt[t], 1u);
                  if (i < (unsigned)LIST_CAP) s_list[t * LIST_CAP + i] = u;
                } else {
                  atomicAdd(&s_hist[t * NB + fine_bin(u, cbn)], 1u);
                }
              }
            }
          }
        }
        __syncthreads();
        if (attempt == 0) {
          const bool overflow = s_cnt[0] > LIST_CAP || s_cnt[1] > LIST_CAP || s_cnt[2] > LIST_CAP || s_cnt[3] > LIST_CAP;
          if (!overflow) {
            fin_exact_select(st, s_list, s_cnt, lst);
            break;
          }
          __syncthreads();
        } else {
          if (lst[1] != 1)
            for (int i = tid; i < NB; i += FUSED_THREADS) s_hist[1 * NB + i] = s_hist[0 * NB + i];
          if (lst[3] != 3)
            for (int i = tid; i < NB; i += FUSED_THREADS) s_hist[3 * NB + i] = s_hist[2 * NB + i];
          __syncthreads();
          fin_fine(st, s_hist, s_scratch);
        }
      }
      fin_conc(C, st);
    }
    __syncthreads();

    // ---- per-image outputs ----
    if (tid == 0) {
      if (status_out) status_out[n] = st->status;
      if (M_src_out)
        for (int i = 0; i < 6; ++i) M_src_out[(size_t)n * 6 + i] = st->M_src[i];
      if (maxC_out) {
        maxC_out[(size_t)n * 2] = st->maxC_src[0];
        maxC_out[(size_t)n * 2 + 1] = st->maxC_src[1];
      }
    }

    // ---- pass G: recompose in place, chunk by chunk, each chunk leaving through a bulk store ----
    if (C.do_apply) {
      const bool ok = st->status == TIAB_STAIN_OK;
      // log2(out_c) = eb_c + sum_ch L_ch * w[ch][c],  w = -W2 / log2(e)  (W2 is defined on OD)
      float w[9], eb[3];
      {
        const float inv_log2e = 0.69314718055994531f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float sq = 0.f;
#pragma unroll
          for (int ch = 0; ch < 3; ++ch) {
            w[ch * 3 + c] = -st->W2[ch * 3 + c] * inv_log2e;
            sq += w[ch * 3 + c];
          }
          eb[c] = LOG2_255 * (1.f - sq);
        }
      }
      for (int c = 0; c < NCHUNK; ++c) {
        if (ok) {
          for (int q = c * chunk_quads + tid; q < (c + 1) * chunk_quads; q += FUSED_THREADS) {
            float L[12];
            quad_logs(s_patch[3 * q], s_patch[3 * q + 1], s_patch[3 * q + 2], L);
            // rgb2od's floor max(OD, 1e-6) (transforms.py:231) only bites for v = 255, but there it decides between
            // 254 and 255 after truncation: L <= log2(255) - 1e-6/ln2
#pragma unroll
            for (int i = 0; i < 12; ++i) L[i] = fminf(L[i], LOG2_255 - 1.4426950e-6f);
            unsigned o[12];
            float fo[12];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
#pragma unroll
              for (int ch = 0; ch < 3; ++ch) {
                const float x = fmaf(L[3 * j], w[ch], fmaf(L[3 * j + 1], w[3 + ch], fmaf(L[3 * j + 2], w[6 + ch], eb[ch])));
                const float v = fminf(mufu_ex2(x), 255.f);
                fo[j * 3 + ch] = v;
                o[j * 3 + ch] = __float2uint_rz(v);
              }
            }
            s_patch[3 * q] = o[0] | (o[1] << 8) | (o[2] << 16) | (o[3] << 24);
            s_patch[3 * q + 1] = o[4] | (o[5] << 8) | (o[6] << 16) | (o[7] << 24);
            s_patch[3 * q + 2] = o[8] | (o[9] << 8) | (o[10] << 16) | (o[11] << 24);
            if (out_f32) {
              float4* f4 = reinterpret_cast<float4*>(out_f32 + ((size_t)n * P + (size_t)q * 4) * 3);
              f4[0] = make_float4(fo[0], fo[1], fo[2], fo[3]);
              f4[1] = make_float4(fo[4], fo[5], fo[6], fo[7]);
              f4[2] = make_float4(fo[8], fo[9], fo[10], fo[11]);
            }
          }
          fence_proxy_async();
        }
        __syncthreads();
        if (tid == 0) {
          bulk_s2g(dst + (size_t)c * chunk_bytes, smem + (size_t)c * chunk_bytes, chunk_bytes);
          bulk_commit();
        }
      }
    }
    __syncthreads();
  }
  if (tid == 0) bulk_wait_all0();
}

// ------------------------------------------------------------------------------------------------
// GENERAL path: one kernel per pass (grid = chunks x images, 256 threads) + block-level finalise
// kernels (grid = images, 1024 threads).
// ------------------------------------------------------------------------------------------------
constexpr int GEN_THREADS = 256;
constexpr int GEN_CHUNK_PX = 16384;

enum {
  PASS_HIST = 0,
  PASS_MOM,
  PASS_PHI_COARSE,
  PASS_PHI_FINE,
  PASS_VAH_NORM,
  PASS_VAH_GC,
  PASS_CONC_COARSE,
  PASS_CONC_FINE,
  PASS_APPLY,
  PASS_MASK
};

struct GenArgs {
  const uint8_t* in;
  uint8_t* out;
  float* out_f32;
  uint8_t* mask_out;
  StainWork* work;
  StainState* state;
  const StainTables* tab;
};

template <int PASS>
__global__ void __launch_bounds__(GEN_THREADS) stain_pass_kernel(const StainConst C, const GenArgs a) {
  __shared__ unsigned sh_hist[(PASS == PASS_HIST) ? 256 : (PASS == PASS_PHI_COARSE ? NB : (PASS == PASS_CONC_COARSE ? 2 * NB : ((PASS == PASS_PHI_FINE || PASS == PASS_CONC_FINE) ? 4 * NB : 1)))];
  __shared__ uint2 sh_lut[768];
  __shared__ float sh_od[256];
  __shared__ double sh_acc[12];
  const int n = blockIdx.y;
  StainState* st = a.state + n;
  StainWork* wk = a.work + n;
  const int tid = threadIdx.x;
  const int P = C.P;
  const uint8_t* base = a.in + (size_t)n * P * 3;
  const int vec = ((reinterpret_cast<uintptr_t>(base) & 3) == 0 && (P & 3) == 0) ? 1 : 0;
  const int px_begin = blockIdx.x * GEN_CHUNK_PX;
  const int px_end = min(P, px_begin + GEN_CHUNK_PX);
  const int status = st->status;
  if (PASS != PASS_APPLY && PASS != PASS_HIST && status != TIAB_STAIN_OK) return;

  constexpr int HW = (PASS == PASS_HIST) ? 256 : (PASS == PASS_PHI_COARSE ? NB : (PASS == PASS_CONC_COARSE ? 2 * NB : ((PASS == PASS_PHI_FINE || PASS == PASS_CONC_FINE) ? 4 * NB : 0)));
  for (int i = tid; i < HW; i += GEN_THREADS) sh_hist[i] = 0u;
  if (PASS != PASS_HIST)
    for (int i = tid; i < 768; i += GEN_THREADS) sh_lut[i] = wk->lut[i];
  for (int i = tid; i < 256; i += GEN_THREADS) sh_od[i] = a.tab->od[i];
  if (tid < 12) sh_acc[tid] = 0.0;
  __syncthreads();
  PixelLut L{sh_lut, sh_od, C.ysum_limit};

  if (PASS == PASS_HIST) {
    for_pixels(base, vec, px_begin, px_end, tid, GEN_THREADS, [&](unsigned r, unsigned g, unsigned b) {
      atomicAdd(&sh_hist[r], 1u);
      atomicAdd(&sh_hist[g], 1u);
      atomicAdd(&sh_hist[b], 1u);
    });
  } else if (PASS == PASS_MOM) {
    float v[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) v[i] = 0.f;
    for_pixels(base, vec, px_begin, px_end, tid, GEN_THREADS, [&](unsigned r, unsigned g, unsigned b) {
      float o0, o1, o2;
      if (L.tissue(r, g, b, o0, o1, o2)) {
        v[0] += 1.f, v[1] += o0, v[2] += o1, v[3] += o2;
        v[4] += o0 * o0, v[5] += o0 * o1, v[6] += o0 * o2, v[7] += o1 * o1, v[8] += o1 * o2, v[9] += o2 * o2;
      }
    });
    block_accumulate<10>(v, sh_acc);
  } else if (PASS == PASS_MASK) {
    float v[1] = {0.f};
    uint8_t* m = a.mask_out + (size_t)n * P;
    for (int p = px_begin + tid; p < px_end; p += GEN_THREADS) {
      float o0, o1, o2;
      bool t = L.tissue(base[3 * p], base[3 * p + 1], base[3 * p + 2], o0, o1, o2);
      m[p] = t ? 1 : 0;
      v[0] += t ? 1.f : 0.f;
    }
    block_accumulate<1>(v, sh_acc);
  } else if (PASS == PASS_PHI_COARSE || PASS == PASS_PHI_FINE) {
    float E[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) E[i] = st->E[i];
    int tb[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) tb[t] = st->tgt_bin[t];
    for_pixels(base, vec, px_begin, px_end, tid, GEN_THREADS, [&](unsigned r, unsigned g, unsigned b) {
      float o0, o1, o2;
      if (L.tissue(r, g, b, o0, o1, o2)) {
        float u = phi_key(E, o0, o1, o2);
        int cb = coarse_bin(u);
        if (PASS == PASS_PHI_COARSE) {
          atomicAdd(&sh_hist[cb], 1u);
        } else {
#pragma unroll
          for (int t = 0; t < 4; ++t)
            if (cb == tb[t]) atomicAdd(&sh_hist[t * NB + fine_bin(u, cb)], 1u);
        }
      }
    });
  } else if (PASS == PASS_VAH_NORM || PASS == PASS_VAH_GC) {
    const int nfull = st->vsteps;
    float v[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) v[i] = 0.f;
    for_pixels(base, vec, px_begin, px_end, tid, GEN_THREADS, [&](unsigned r, unsigned g, unsigned b) {
      float o0, o1, o2;
      if (L.tissue(r, g, b, o0, o1, o2)) {
        float d0, d1, pend = 0.f;
        vah_replay(st, nfull, PASS == PASS_VAH_NORM, o0, o1, o2, d0, d1, pend);
        if (PASS == PASS_VAH_NORM) {
          v[0] += pend * pend;
        } else {
          v[0] += d0 * d0, v[1] += d0 * d1, v[2] += d1 * d1;
          v[3] += d0 * o0, v[4] += d0 * o1, v[5] += d0 * o2, v[6] += d1 * o0, v[7] += d1 * o1, v[8] += d1 * o2;
        }
      }
    });
    block_accumulate<9>(v, sh_acc);
  } else if (PASS == PASS_CONC_COARSE || PASS == PASS_CONC_FINE) {
    float A[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) A[i] = st->A[i];
    const float clo0 = st->c_lo[0], clo1 = st->c_lo[1], cs0 = st->c_scale[0], cs1 = st->c_scale[1];
    int tb[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) tb[t] = st->tgt_bin[t];
    for_pixels(base, vec, px_begin, px_end, tid, GEN_THREADS, [&](unsigned r, unsigned g, unsigned b) {
      float o0 = sh_od[r], o1 = sh_od[g], o2 = sh_od[b];
      float u0 = (o0 * A[0] + o1 * A[2] + o2 * A[4] - clo0) * cs0;
      float u1 = (o0 * A[1] + o1 * A[3] + o2 * A[5] - clo1) * cs1;
      int cb0 = coarse_bin(u0), cb1 = coarse_bin(u1);
      if (PASS == PASS_CONC_COARSE) {
        atomicAdd(&sh_hist[cb0], 1u);
        atomicAdd(&sh_hist[NB + cb1], 1u);
      } else {
        if (cb0 == tb[0]) atomicAdd(&sh_hist[0 * NB + fine_bin(u0, cb0)], 1u);
        if (cb0 == tb[1]) atomicAdd(&sh_hist[1 * NB + fine_bin(u0, cb0)], 1u);
        if (cb1 == tb[2]) atomicAdd(&sh_hist[2 * NB + fine_bin(u1, cb1)], 1u);
        if (cb1 == tb[3]) atomicAdd(&sh_hist[3 * NB + fine_bin(u1, cb1)], 1u);
      }
    });
  } else if (PASS == PASS_APPLY) {
    uint8_t* dst = a.out + (size_t)n * P * 3;
    const bool ok = status == TIAB_STAIN_OK;
    float W[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) W[i] = st->W2[i];
    const float l255 = 7.9943534368588578f;
    for (int p = px_begin + tid; p < px_end; p += GEN_THREADS) {
      unsigned r = base[3 * p], g = base[3 * p + 1], b = base[3 * p + 2];
      if (!ok) {
        dst[3 * p] = r, dst[3 * p + 1] = g, dst[3 * p + 2] = b;
        continue;
      }
      float o0 = sh_od[r], o1 = sh_od[g], o2 = sh_od[b];
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        float e = fmaf(o0, W[ch], fmaf(o1, W[3 + ch], fmaf(o2, W[6 + ch], l255)));
        float v = fminf(exp2f(e), 255.f);
        dst[3 * p + ch] = (uint8_t)__float2uint_rz(v);
        if (a.out_f32) a.out_f32[((size_t)n * P + p) * 3 + ch] = v;
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < HW; i += GEN_THREADS)
    if (sh_hist[i]) atomicAdd(&wk->hist[i], sh_hist[i]);
  if (PASS == PASS_MOM || PASS == PASS_VAH_NORM || PASS == PASS_VAH_GC || PASS == PASS_MASK)
    if (tid < 12 && sh_acc[tid] != 0.0) atomicAdd(&st->acc[tid], sh_acc[tid]);
}

enum { FIN_INIT = 0, FIN_HIST, FIN_MOM, FIN_PHI_COARSE, FIN_PHI_FINE, FIN_VAH_PREP, FIN_VAH_NORM, FIN_VAH_GC, FIN_FIXED, FIN_CONC_PREP, FIN_CONC_COARSE, FIN_CONC_FINE, FIN_OUTPUT, FIN_MASK };

template <int STEP>
__global__ void __launch_bounds__(FUSED_THREADS) stain_fin_kernel(const StainConst C, const GenArgs a, double* M_src_out,
                                                                  double* maxC_out, int* status_out) {
  __shared__ unsigned scratch[64];
  const int n = blockIdx.x;
  StainState* st = a.state + n;
  StainWork* wk = a.work + n;
  const int tid = threadIdx.x;
  auto zero_hist = [&]() {
    for (int i = tid; i < HIST_WORDS; i += FUSED_THREADS) wk->hist[i] = 0u;
  };
  auto zero_acc = [&]() {
    if (tid < 12) st->acc[tid] = 0.0;
  };
  if (STEP == FIN_INIT) {
    if (tid == 0) {
      st->status = TIAB_STAIN_OK;
      st->n_tissue = 0;
      st->vsteps = 0;
      st->vah.done = 0;
      st->vah.nsteps = 0;
      for (int i = 0; i < 6; ++i) st->M_src[i] = 0.0;
      st->maxC_src[0] = st->maxC_src[1] = 0.0;
    }
    zero_hist();
    zero_acc();
    return;
  }
  if (STEP == FIN_HIST) {
    fin_hist(C, a.tab, st, wk->hist, wk->lut);
    zero_acc();
    return;
  }
  if (STEP == FIN_MASK) {
    if (tid == 0) {
      st->n_tissue = (unsigned)(st->acc[0] + 0.5);
      if (st->n_tissue == 0) st->status = TIAB_STAIN_EMPTY_TISSUE_MASK;
      if (status_out) status_out[n] = st->status;
    }
    return;
  }
  if (STEP == FIN_OUTPUT) {
    if (tid == 0) {
      if (status_out) status_out[n] = st->status;
      if (M_src_out)
        for (int i = 0; i < 6; ++i) M_src_out[(size_t)n * 6 + i] = st->M_src[i];
      if (maxC_out) maxC_out[(size_t)n * 2] = st->maxC_src[0], maxC_out[(size_t)n * 2 + 1] = st->maxC_src[1];
    }
    return;
  }
  if (STEP == FIN_FIXED) {
    if (tid == 0) set_source_matrix(st, C.M_fixed);
    return;
  }
  if (st->status != TIAB_STAIN_OK) return;
  if (STEP == FIN_MOM) {
    fin_moments(C, st);
    zero_hist();
    if (tid == 0 && st->status == TIAB_STAIN_OK && C.method == M_MACENKO) {
      set_rank_targets(st, 0, 0, st->n_tissue, C.q_phi_lo);
      set_rank_targets(st, 2, 1, st->n_tissue, C.q_phi_hi);
    }
  } else if (STEP == FIN_PHI_COARSE) {
    fin_coarse(st, wk->hist, 0, 0, scratch);
    zero_hist();
  } else if (STEP == FIN_PHI_FINE) {
    fin_fine(st, wk->hist, scratch);
    fin_phi(C, st);
  } else if (STEP == FIN_VAH_PREP) {
    if (!st->vah.done) vah_prepare_step(C, st);
  } else if (STEP == FIN_VAH_NORM) {
    if (!st->vah.done) vah_finish_norm(st);
  } else if (STEP == FIN_VAH_GC) {
    if (!st->vah.done) vah_finish_gc(C, st);
    zero_acc();
  } else if (STEP == FIN_CONC_PREP) {
    zero_hist();
    if (tid == 0) {
      set_rank_targets(st, 0, 0, (unsigned)C.P, C.q_conc);
      set_rank_targets(st, 2, 1, (unsigned)C.P, C.q_conc);
    }
  } else if (STEP == FIN_CONC_COARSE) {
    fin_coarse(st, wk->hist, 0, 1, scratch);
    zero_hist();
  } else if (STEP == FIN_CONC_FINE) {
    fin_fine(st, wk->hist, scratch);
    fin_conc(C, st);
  }
}

// StainNormalizer.get_concentrations (stainnorm.py:49-66): C = od @ pinv(M), float32 per pixel.
__global__ void __launch_bounds__(256) stain_conc_kernel(const StainTables* __restrict__ tab,
                                                         const uint8_t* __restrict__ in, long long npix, float a0,
                                                         float a1, float a2, float a3, float a4, float a5,
                                                         float* __restrict__ out) {
  __shared__ float sh_od[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) sh_od[i] = tab->od[i];
  __syncthreads();
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += (long long)gridDim.x * blockDim.x) {
    float o0 = sh_od[in[3 * p]], o1 = sh_od[in[3 * p + 1]], o2 = sh_od[in[3 * p + 2]];
    out[2 * p] = o0 * a0 + o1 * a2 + o2 * a4;
    out[2 * p + 1] = o0 * a1 + o1 * a3 + o2 * a5;
  }
}

}  // namespace tiab

// ================================================================================================
// host side
// ================================================================================================
using namespace tiab;

static void host_tables(StainTables* t) {
  for (int v = 0; v < 256; ++v) {
    t->od[v] = (float)od_of_level(v);
    t->tabg[v] = srgb_gamma_2040(v);
  }
}

static int ensure_bytes(tiab_ctx* h, void** p, size_t* cur, size_t need) {
  if (*cur >= need) return TIAB_OK;
  if (*p) cudaFree(*p);
  *p = nullptr;
  *cur = 0;
  TIAB_CUDA_OK(h, cudaMalloc(p, need));
  *cur = need;
  return TIAB_OK;
}

static int stain_setup(tiab_ctx* h) {
  if (!h->stain_lut) {
    StainTables t;
    host_tables(&t);
    TIAB_CUDA_OK(h, cudaMalloc(&h->stain_lut, sizeof(StainTables)));
    TIAB_CUDA_OK(h, cudaMemcpy(h->stain_lut, &t, sizeof(t), cudaMemcpyHostToDevice));
    TIAB_CUDA_OK(h, cudaFuncSetAttribute(stain_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         FusedSmem::TOTAL));
  }
  return TIAB_OK;
}

static int make_const(tiab_ctx* h, const tiab_stain_params_t* p, int H, int W, const double* Mt, const double* mc,
                      int do_apply, StainConst* C) {
  if (!p || H <= 0 || W <= 0 || (long long)H * W > (1LL << 30)) {
    h->err = "invalid stain arguments";
    return TIAB_E_INVALID;
  }
  if (p->method < 0 || p->method > 3) {
    h->err = "invalid stain method";
    return TIAB_E_INVALID;
  }
  memset(C, 0, sizeof(*C));
  C->method = p->method;
  C->H = H;
  C->W = W;
  C->P = H * W;
  C->ysum_limit = ((p->y_threshold + 1) << 12) - 2048;
  C->do_apply = do_apply;
  C->q_lo = 2.0 / 100.0;
  C->q_hi = 98.0 / 100.0;
  C->q_phi_lo = (100.0 - p->angular_percentile) / 100.0;
  C->q_phi_hi = p->angular_percentile / 100.0;
  C->q_conc = 99.0 / 100.0;
  C->reg = p->regularizer;
  for (int i = 0; i < 6; ++i) C->M_fixed[i] = p->fixed_matrix[i];
  if (Mt)
    for (int i = 0; i < 6; ++i) C->M_target[i] = Mt[i];
  if (mc) C->maxC_target[0] = mc[0], C->maxC_target[1] = mc[1];
  return TIAB_OK;
}

static bool fused_ok(const StainConst& C, const void* in, const void* out) {
  long long bytes = (long long)C.P * 3;
  return bytes <= PATCH_BYTES_MAX && (C.P % 128) == 0 && ((uintptr_t)in & 15) == 0 && ((uintptr_t)out & 15) == 0;
}

#define LAUNCH_CHECK(h)                                  \
  do {                                                   \
    cudaError_t _e = cudaGetLastError();                 \
    if (_e != cudaSuccess) {                             \
      (h)->err = std::string("kernel launch: ") + cudaGetErrorString(_e); \
      return TIAB_E_CUDA;                                \
    }                                                    \
    (h)->last_launches++;                                \
  } while (0)

// the multi-pass path over a slab of `n` images
static int run_general(tiab_ctx* h, const StainConst& C, const uint8_t* in, uint8_t* out, float* out_f32, int n,
                       double* M_src, double* maxC, int* status, cudaStream_t s) {
  int rc;
  if ((rc = ensure_bytes(h, &h->stain_work, &h->stain_work_bytes, sizeof(StainWork) * (size_t)n))) return rc;
  if ((rc = ensure_bytes(h, &h->stain_state, &h->stain_state_bytes, sizeof(StainState) * (size_t)n))) return rc;
  GenArgs a{in, out, out_f32, nullptr, (StainWork*)h->stain_work, (StainState*)h->stain_state,
            (const StainTables*)h->stain_lut};
  dim3 pg((C.P + GEN_CHUNK_PX - 1) / GEN_CHUNK_PX, n), pb(GEN_THREADS);
  dim3 fg(n), fb(FUSED_THREADS);
#define PASS(K) stain_pass_kernel<K><<<pg, pb, 0, s>>>(C, a); LAUNCH_CHECK(h)
#define FIN(K) stain_fin_kernel<K><<<fg, fb, 0, s>>>(C, a, M_src, maxC, status); LAUNCH_CHECK(h)
  FIN(FIN_INIT);
  if (C.method == M_MACENKO || C.method == M_VAHADANE) {
    PASS(PASS_HIST);
    FIN(FIN_HIST);
    PASS(PASS_MOM);
    FIN(FIN_MOM);
    if (C.method == M_MACENKO) {
      PASS(PASS_PHI_COARSE);
      FIN(FIN_PHI_COARSE);
      PASS(PASS_PHI_FINE);
      FIN(FIN_PHI_FINE);
    } else {
      for (int it = 0; it < 2; ++it) {
        for (int half = 0; half < 2; ++half) {
          FIN(FIN_VAH_PREP);
          PASS(PASS_VAH_NORM);
          FIN(FIN_VAH_NORM);
        }
        PASS(PASS_VAH_GC);
        FIN(FIN_VAH_GC);
      }
    }
  } else {
    FIN(FIN_FIXED);
  }
  FIN(FIN_CONC_PREP);
  PASS(PASS_CONC_COARSE);
  FIN(FIN_CONC_COARSE);
  PASS(PASS_CONC_FINE);
  FIN(FIN_CONC_FINE);
  FIN(FIN_OUTPUT);
  if (C.do_apply) {
    PASS(PASS_APPLY);
  }
#undef PASS
#undef FIN
  return TIAB_OK;
}

static int run_stain(tiab_ctx* h, const StainConst& C, const uint8_t* in, uint8_t* out, float* out_f32, int N,
                     double* M_src, double* maxC, int* status, cudaStream_t s) {
  int rc = stain_setup(h);
  if (rc) return rc;
  h->last_launches = 0;
  bool fused = fused_ok(C, in, out ? out : in);
  if (h->stain_path == 2) fused = false;
  if (h->stain_path == 1 && !fused) {
    h->err = "fused stain path requested but the shape/alignment does not allow it";
    return TIAB_E_UNSUPPORTED;
  }
  if (fused) {
    int grid = std::min(N, h->sm_count);
    stain_fused_kernel<<<grid, FUSED_THREADS, FusedSmem::TOTAL, s>>>(C, (const StainTables*)h->stain_lut, in,
                                                                    out ? out : const_cast<uint8_t*>(in), N, M_src,
                                                                    maxC, status, out_f32);
    LAUNCH_CHECK(h);
    return TIAB_OK;
  }
  // general: slabs of images so that the scratch stays bounded (<= 64 MB)
  const int slab = std::max(1, std::min(N, 2048));
  for (int n0 = 0; n0 < N; n0 += slab) {
    int n = std::min(slab, N - n0);
    size_t off = (size_t)n0 * C.P * 3;
    rc = run_general(h, C, in + off, out ? out + off : nullptr, out_f32 ? out_f32 + off : nullptr, n,
                     M_src ? M_src + (size_t)n0 * 6 : nullptr, maxC ? maxC + (size_t)n0 * 2 : nullptr,
                     status ? status + n0 : nullptr, s);
    if (rc) return rc;
  }
  return TIAB_OK;
}

extern "C" {

int tiab_stain_set_path(tiab_handle_t h, int path) {
  if (!h || path < 0 || path > 2) return TIAB_E_INVALID;
  h->stain_path = path;
  return TIAB_OK;
}
int tiab_last_launch_count(tiab_handle_t h) { return h ? h->last_launches : 0; }

int tiab_stain_transform_batch(tiab_handle_t h, const tiab_stain_params_t* p, const uint8_t* in_dev, uint8_t* out_dev,
                               int N, int H, int W, const double M_target[6], const double maxC_target[2],
                               double* M_src_dev, double* maxC_src_dev, int* status_dev, float* out_f32_dev,
                               void* stream) {
  if (!h) return TIAB_E_INVALID;
  if (!in_dev || !out_dev || N <= 0 || !M_target || !maxC_target) {
    h->err = "tiab_stain_transform_batch: null/empty argument";
    return TIAB_E_INVALID;
  }
  StainConst C;
  int rc = make_const(h, p, H, W, M_target, maxC_target, 1, &C);
  if (rc) return rc;
  TIAB_ON_DEVICE(h);
  return run_stain(h, C, in_dev, out_dev, out_f32_dev, N, M_src_dev, maxC_src_dev, status_dev, (cudaStream_t)stream);
}

static int ensure_small(tiab_ctx* h, int N) {
  size_t need = (size_t)N * (8 * sizeof(double) + sizeof(int));
  if (h->small_bytes >= need) return TIAB_OK;
  if (h->dev_small) cudaFree(h->dev_small);
  if (h->pin_small) cudaFreeHost(h->pin_small);
  h->dev_small = h->pin_small = nullptr;
  h->small_bytes = 0;
  TIAB_CUDA_OK(h, cudaMalloc(&h->dev_small, need));
  TIAB_CUDA_OK(h, cudaMallocHost(&h->pin_small, need));
  h->small_bytes = need;
  return TIAB_OK;
}

int tiab_stain_fit(tiab_handle_t h, const tiab_stain_params_t* p, const uint8_t* target_dev, int H, int W,
                   double M_target_host[6], double maxC_target_host[2], int* status_host, void* stream) {
  if (!h) return TIAB_E_INVALID;
  if (!target_dev || !M_target_host || !maxC_target_host) {
    h->err = "tiab_stain_fit: null argument";
    return TIAB_E_INVALID;
  }
  StainConst C;
  int rc = make_const(h, p, H, W, nullptr, nullptr, 0, &C);
  if (rc) return rc;
  TIAB_ON_DEVICE(h);
  if ((rc = ensure_small(h, 1))) return rc;
  double* dM = (double*)h->dev_small;
  double* dC = dM + 6;
  int* dS = (int*)(dC + 2);
  cudaStream_t s = (cudaStream_t)stream;
  rc = run_stain(h, C, target_dev, nullptr, nullptr, 1, dM, dC, dS, s);
  if (rc) return rc;
  double* hp = (double*)h->pin_small;
  TIAB_CUDA_OK(h, cudaMemcpyAsync(hp, h->dev_small, 8 * sizeof(double) + sizeof(int), cudaMemcpyDeviceToHost, s));
  TIAB_CUDA_OK(h, cudaStreamSynchronize(s));
  for (int i = 0; i < 6; ++i) M_target_host[i] = hp[i];
  maxC_target_host[0] = hp[6];
  maxC_target_host[1] = hp[7];
  if (status_host) *status_host = *(int*)(hp + 8);
  return TIAB_OK;
}

static int ensure_staging(tiab_ctx* h, size_t bytes) {
  if (h->pin_bytes < bytes) {
    if (h->pin_in) cudaFreeHost(h->pin_in);
    if (h->pin_out) cudaFreeHost(h->pin_out);
    h->pin_in = h->pin_out = nullptr;
    h->pin_bytes = 0;
    TIAB_CUDA_OK(h, cudaMallocHost(&h->pin_in, bytes));
    TIAB_CUDA_OK(h, cudaMallocHost(&h->pin_out, bytes));
    h->pin_bytes = bytes;
  }
  if (h->dev_bytes < bytes) {
    if (h->dev_in) cudaFree(h->dev_in);
    if (h->dev_out) cudaFree(h->dev_out);
    h->dev_in = h->dev_out = nullptr;
    h->dev_bytes = 0;
    TIAB_CUDA_OK(h, cudaMalloc(&h->dev_in, bytes));
    TIAB_CUDA_OK(h, cudaMalloc(&h->dev_out, bytes));
    h->dev_bytes = bytes;
  }
  return TIAB_OK;
}

int tiab_stain_fit_host(tiab_handle_t h, const tiab_stain_params_t* p, const uint8_t* target_host, int H, int W,
                        double M_target_host[6], double maxC_target_host[2], int* status_host) {
  if (!h) return TIAB_E_INVALID;
  if (!target_host || H <= 0 || W <= 0) {
    h->err = "tiab_stain_fit_host: null/empty argument";
    return TIAB_E_INVALID;
  }
  TIAB_ON_DEVICE(h);
  size_t bytes = (size_t)H * W * 3;
  int rc = ensure_staging(h, bytes);
  if (rc) return rc;
  memcpy(h->pin_in, target_host, bytes);
  TIAB_CUDA_OK(h, cudaMemcpyAsync(h->dev_in, h->pin_in, bytes, cudaMemcpyHostToDevice, h->own_stream));
  return tiab_stain_fit(h, p, (const uint8_t*)h->dev_in, H, W, M_target_host, maxC_target_host, status_host,
                        h->own_stream);
}

// memcpy split over a few host threads: one core moves ~10 GB/s, which is less than the PCIe link of a B200 host (the r1 host path —
// one memcpy into a pinned buffer, one blocking copy, one memcpy out — ran at 5 GB/s end to end)
static void parallel_memcpy(void* dst, const void* src, size_t n, int threads) {
  if (threads <= 1 || n < ((size_t)8 << 20)) {
    memcpy(dst, src, n);
    return;
  }
  std::vector<std::thread> pool;
  const size_t per = ((n / threads) + 4095) & ~(size_t)4095;
  for (int t = 1; t < threads; ++t) {
    const size_t o = std::min(n, per * t), e = std::min(n, per * (t + 1));
    if (e > o) pool.emplace_back([=] { memcpy((uint8_t*)dst + o, (const uint8_t*)src + o, e - o); });
  }
  memcpy(dst, src, std::min(n, per));
  for (std::thread& th : pool) th.join();
}

// Host-buffer entry point (what a per-patch `preproc_func` user or the CLI gets, tools/stainnorm.py:89-113): the batch streams through
// THREE pinned / device slot pairs in chunks — while chunk c runs on the GPU, chunk c + 1 is staged and copied up and chunk c - 1 is
// copied down and handed back — so the call is bounded by the slower of the host memcpys and one PCIe direction, not by their sum.
int tiab_stain_transform_host(tiab_handle_t h, const tiab_stain_params_t* p, const uint8_t* in_host, uint8_t* out_host,
                              int N, int H, int W, const double M_target[6], const double maxC_target[2],
                              double* M_src_host, double* maxC_src_host, int* status_host) {
  if (!h) return TIAB_E_INVALID;
  if (!in_host || !out_host || N <= 0 || H <= 0 || W <= 0) {
    h->err = "tiab_stain_transform_host: null/empty argument";
    return TIAB_E_INVALID;
  }
  TIAB_ON_DEVICE(h);
  constexpr int SLOTS = 3;
  const size_t tile = (size_t)H * W * 3;
  // chunk: ~32 MB of tiles, at least one SM-count worth for the fused kernel when the batch allows it
  int chunk = (int)std::max<size_t>(1, ((size_t)32 << 20) / tile);
  chunk = std::min(chunk, N);
  if (N > SLOTS && chunk * 2 > N) chunk = (N + 1) / 2;
  const size_t chunk_bytes = (size_t)chunk * tile;
  int rc = ensure_staging(h, chunk_bytes * SLOTS);
  if (rc) return rc;
  if ((rc = ensure_small(h, N))) return rc;
  if (!h->copy_stream) TIAB_CUDA_OK(h, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
  cudaStream_t s = h->own_stream;
  double* dM = (double*)h->dev_small;
  double* dC = dM + (size_t)N * 6;
  int* dS = (int*)(dC + (size_t)N * 2);
  const int nchunks = (N + chunk - 1) / chunk;
  const int threads = (int)std::min<unsigned>(8u, std::max(1u, std::thread::hardware_concurrency() / 4));
  cudaEvent_t up[SLOTS], done[SLOTS], down[SLOTS];
  for (int i = 0; i < SLOTS; ++i) {
    TIAB_CUDA_OK(h, cudaEventCreateWithFlags(&up[i], cudaEventDisableTiming));
    TIAB_CUDA_OK(h, cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming));
    TIAB_CUDA_OK(h, cudaEventCreateWithFlags(&down[i], cudaEventDisableTiming));
  }
  auto cleanup = [&]() {
    for (int i = 0; i < SLOTS; ++i) cudaEventDestroy(up[i]), cudaEventDestroy(done[i]), cudaEventDestroy(down[i]);
  };
  auto slot_ptr = [&](void* base, int slot) { return (uint8_t*)base + (size_t)slot * chunk_bytes; };
  auto finish = [&](int c) -> int {   // chunk c's result has arrived in its pinned slot: hand it back
    const int slot = c % SLOTS, n0 = c * chunk, n = std::min(chunk, N - n0);
    if (cudaEventSynchronize(down[slot]) != cudaSuccess) return TIAB_E_CUDA;
    parallel_memcpy(out_host + (size_t)n0 * tile, slot_ptr(h->pin_out, slot), (size_t)n * tile, threads);
    return TIAB_OK;
  };
  for (int c = 0; c < nchunks; ++c) {
    const int slot = c % SLOTS, n0 = c * chunk, n = std::min(chunk, N - n0);
    if (c >= SLOTS && (rc = finish(c - SLOTS))) {      // frees this slot's pinned output (and, by stream order, its device buffers)
      cleanup();
      h->err = "tiab_stain_transform_host: device -> host copy failed";
      return rc;
    }
    parallel_memcpy(slot_ptr(h->pin_in, slot), in_host + (size_t)n0 * tile, (size_t)n * tile, threads);
    // upload on the copy stream (overlaps the previous chunk's kernel and download), compute + download on the handle's stream
    if (c >= SLOTS) cudaStreamWaitEvent(h->copy_stream, done[slot], 0);   // the kernel that read this device slot has finished
    cudaMemcpyAsync(slot_ptr(h->dev_in, slot), slot_ptr(h->pin_in, slot), (size_t)n * tile, cudaMemcpyHostToDevice, h->copy_stream);
    cudaEventRecord(up[slot], h->copy_stream);
    cudaStreamWaitEvent(s, up[slot], 0);
    rc = tiab_stain_transform_batch(h, p, slot_ptr(h->dev_in, slot), slot_ptr(h->dev_out, slot), n, H, W, M_target, maxC_target,
                                    dM + (size_t)n0 * 6, dC + (size_t)n0 * 2, dS + n0, nullptr, s);
    if (rc) {
      cudaStreamSynchronize(s), cudaStreamSynchronize(h->copy_stream);
      cleanup();
      return rc;
    }
    cudaEventRecord(done[slot], s);
    cudaMemcpyAsync(slot_ptr(h->pin_out, slot), slot_ptr(h->dev_out, slot), (size_t)n * tile, cudaMemcpyDeviceToHost, s);
    cudaEventRecord(down[slot], s);
  }
  for (int c = std::max(0, nchunks - SLOTS); c < nchunks; ++c)
    if ((rc = finish(c))) {
      cleanup();
      h->err = "tiab_stain_transform_host: device -> host copy failed";
      return rc;
    }
  size_t small = (size_t)N * (8 * sizeof(double) + sizeof(int));
  cudaError_t ce = cudaMemcpyAsync(h->pin_small, h->dev_small, small, cudaMemcpyDeviceToHost, s);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(s);
  cleanup();
  TIAB_CUDA_OK(h, ce);
  TIAB_CUDA_OK(h, cudaGetLastError());
  const double* hp = (const double*)h->pin_small;
  if (M_src_host) memcpy(M_src_host, hp, (size_t)N * 6 * sizeof(double));
  if (maxC_src_host) memcpy(maxC_src_host, hp + (size_t)N * 6, (size_t)N * 2 * sizeof(double));
  if (status_host) memcpy(status_host, hp + (size_t)N * 8, (size_t)N * sizeof(int));
  return TIAB_OK;
}

int tiab_stain_tissue_mask(tiab_handle_t h, const tiab_stain_params_t* p, const uint8_t* in_dev, uint8_t* mask_dev,
                           int N, int H, int W, int* status_dev, void* stream) {
  if (!h) return TIAB_E_INVALID;
  if (!in_dev || !mask_dev || N <= 0) {
    h->err = "tiab_stain_tissue_mask: null/empty argument";
    return TIAB_E_INVALID;
  }
  StainConst C;
  int rc = make_const(h, p, H, W, nullptr, nullptr, 0, &C);
  if (rc) return rc;
  TIAB_ON_DEVICE(h);
  if ((rc = stain_setup(h))) return rc;
  h->last_launches = 0;
  cudaStream_t s = (cudaStream_t)stream;
  const int slab = std::min(N, 2048);
  for (int n0 = 0; n0 < N; n0 += slab) {
    int n = std::min(slab, N - n0);
    if ((rc = ensure_bytes(h, &h->stain_work, &h->stain_work_bytes, sizeof(StainWork) * (size_t)n))) return rc;
    if ((rc = ensure_bytes(h, &h->stain_state, &h->stain_state_bytes, sizeof(StainState) * (size_t)n))) return rc;
    GenArgs a{in_dev + (size_t)n0 * C.P * 3, nullptr, nullptr, mask_dev + (size_t)n0 * C.P, (StainWork*)h->stain_work,
              (StainState*)h->stain_state, (const StainTables*)h->stain_lut};
    dim3 pg((C.P + GEN_CHUNK_PX - 1) / GEN_CHUNK_PX, n), pb(GEN_THREADS), fg(n), fb(FUSED_THREADS);
    stain_fin_kernel<FIN_INIT><<<fg, fb, 0, s>>>(C, a, nullptr, nullptr, nullptr);
    LAUNCH_CHECK(h);
    stain_pass_kernel<PASS_HIST><<<pg, pb, 0, s>>>(C, a);
    LAUNCH_CHECK(h);
    stain_fin_kernel<FIN_HIST><<<fg, fb, 0, s>>>(C, a, nullptr, nullptr, nullptr);
    LAUNCH_CHECK(h);
    stain_pass_kernel<PASS_MASK><<<pg, pb, 0, s>>>(C, a);
    LAUNCH_CHECK(h);
    stain_fin_kernel<FIN_MASK><<<fg, fb, 0, s>>>(C, a, nullptr, nullptr, status_dev ? status_dev + n0 : nullptr);
    LAUNCH_CHECK(h);
  }
  return TIAB_OK;
}

int tiab_stain_concentrations(tiab_handle_t h, const uint8_t* in_dev, int N, int H, int W, const double M[6],
                              float* conc_dev, void* stream) {
  if (!h) return TIAB_E_INVALID;
  if (!in_dev || !conc_dev || !M || N <= 0 || H <= 0 || W <= 0) {
    h->err = "tiab_stain_concentrations: null/empty argument";
    return TIAB_E_INVALID;
  }
  double A[6];
  if (!pinv23(M, A)) {
    h->err = "tiab_stain_concentrations: singular stain matrix";
    return TIAB_E_INVALID;
  }
  TIAB_ON_DEVICE(h);
  int rc = stain_setup(h);
  if (rc) return rc;
  long long npix = (long long)N * H * W;
  int grid = (int)std::min<long long>((npix + 255) / 256, (long long)h->sm_count * 16);
  h->last_launches = 0;
  stain_conc_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const StainTables*)h->stain_lut, in_dev, npix, (float)A[0],
                                                            (float)A[1], (float)A[2], (float)A[3], (float)A[4],
                                                            (float)A[5], conc_dev);
  LAUNCH_CHECK(h);
  return TIAB_OK;
}

// ---- host-callable scalar pieces for the CPU unit tests ----
int tiab_dbg_percentile_u8(const unsigned* hist256, double q, double* out) {
  unsigned n = 0;
  for (int i = 0; i < 256; ++i) n += hist256[i];
  if (!n) return TIAB_E_INVALID;
  *out = percentile_u8(hist256, n, q);
  return TIAB_OK;
}
int tiab_dbg_contrast_lut(const unsigned* hist256, unsigned n, uint8_t lut_out[256]) {
  double lo, hi;
  contrast_range(hist256, n, 2.0 / 100.0, 98.0 / 100.0, &lo, &hi);
  for (int v = 0; v < 256; ++v) lut_out[v] = (uint8_t)contrast_value(v, lo, hi);
  return TIAB_OK;
}
int tiab_dbg_eigh3(const double cov[9], double evals[3], double evecs[9]) {
  eigh3(cov, evals, evecs);
  return TIAB_OK;
}
int tiab_dbg_lasso2(const double gram[4], const double c[2], double reg, double n, double out[2]) {
  lasso2(gram, c, reg, n, out);
  return TIAB_OK;
}
int tiab_dbg_macenko_from_angles(const double evec[6], double phi_min, double phi_max, double M_out[6]) {
  macenko_from_angles(evec, phi_min, phi_max, M_out);
  return TIAB_OK;
}
int tiab_dbg_tables(unsigned tabg[256], float od[256]) {
  StainTables t;
  host_tables(&t);
  memcpy(tabg, t.tabg, sizeof(t.tabg));
  memcpy(od, t.od, sizeof(t.od));
  return TIAB_OK;
}

}  // extern "C"
