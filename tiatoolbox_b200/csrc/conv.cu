// conv.cu — B200 (sm_100a) convolution engine for the patch CNNs (resnet18-kather100k,
// fcn-tissue_mask, hovernet_fast-pannuke).
//
// One tcgen05 kernel does every dense contraction of the three networks as a *shifted GEMM*:
//
//     D[p, co] = sum_t  sum_ci  X[p + off_t, ci] * W[co, t, ci]            (fp16 x fp16 -> fp32)
//
// where X is an NHWC activation tensor viewed as a [rows, C] matrix, p runs over the *input* grid
// ("virtual" output pixels: positions whose receptive field leaves the valid window are computed
// and discarded by the epilogue), and off_t = kh*W_in + kw is a plain row offset per filter tap.
// A stride-1 KxK convolution on a (pre-padded) NHWC buffer therefore needs no im2col and no 4-D
// boxes: every A tile is one 2-D TMA load of 128 consecutive rows x 64 channels at row p0 + off_t
// (out-of-range rows are zero-filled by TMA).  1x1 convs are the T = 1 case; strided convs and the
// 3-channel stems go through an explicit im2col (few layers) and then the same kernel.
//
// Pipeline (CUTLASS-free, inline PTX): warp 0 = TMA producer, warp 1 = single-thread tcgen05.mma
// issuer (UMMA 128 x BLOCK_N x 16, kind::f16, accumulators in TMEM, double buffered), warp 2 =
// TMEM allocator, warps 4-7 = epilogue (tcgen05.ld -> bias / residual / ReLU / second
// BN+ReLU output -> NHWC stores with window + destination remapping).  Persistent CTAs, static
// tile scheduler, S-stage smem ring with full/empty mbarriers.
//
// Reference layers this replaces: torch.nn.Conv2d/BatchNorm2d/ReLU inside
// tiatoolbox/models/architecture/hovernet.py:61-262,330-454, unet.py:24-418, vanilla.py:256-333
// (CPU: oneDNN, CUDA: cuDNN library calls).
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <string>

#include "../../include/tiab200.h"
#include "ctx.h"

namespace tiab {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // fp16: 128 bytes = one SWIZZLE_128B row
constexpr int UMMA_K = 16;
constexpr int CONV_THREADS = 384;  // warps 0-3: TMA / MMA / TMEM alloc / idle ; warps 4-11: epilogue (2 per TMEM lane quadrant)
constexpr int MAX_STAGES = 8;
constexpr int CONV_THREADS_T = 512;  // transform mode: warps 4-7 epilogue, warps 8-15 = two groups of 4 operand-transform warps

struct ConvParams {
  int M, N;
  int taps, kchunks;
  int tap_off[TIAB_MAX_TAPS];
  int tap_coff[TIAB_MAX_TAPS];   // channel offset of each tap's operand inside the x rows (space-to-depth inputs), else 0
  int s2d_c;                     // > 0: out16 is stored space-to-depth (see tiab_conv_desc_t.s2d_c)
  int Hin, Win;           // virtual grid (rows per image = Hin*Win)
  int y0, y1, x0, x1;     // valid output window on the grid
  int Hd, Wd, dy, dx;     // destination row = (img*Hd + y + dy)*Wd + x + dx
  int Hr, Wr, ry, rx;     // residual row    = (img*Hr + y + ry)*Wr + x + rx
  __half* out16;
  int ld16, coff16;
  float* out32;
  int ld32, coff32;
  __half* out16b;
  int ld16b, coff16b;
  const float* scale2;
  const float* shift2;
  const float* bias;
  const float* oscale;    // per-output-channel factor applied to the accumulator before the bias (weights are stored pre-scaled), or null
  const float* res32;
  const __half* res16;
  int ldres, coffres;
  int relu;
  // A-operand "transform" mode (AMODE == 1): A tiles are produced by 4 warps from a raw tensor instead of TMA:
  //   a = relu?(x * pre_scale[c] + pre_shift[c]), x fp32 (residual stream) or fp16 hi(+lo); written as hi/lo swizzled tiles.
  // This is the pre-activation BN+ReLU of hovernet.py:84-86 / :170-172 fused into the consumer GEMM.
  const void* a_src;
  long long a_rows;
  int a_ld, a_lo, a_is_f32, a_cin;
  const float* pre_scale;
  const float* pre_shift;
  int pre_relu;
  int nseg;               // 1: fp16 operands ; 3: split-fp16 (A_hi*W_hi + A_lo*W_hi + A_hi*W_lo ~ fp32 operands)
  int lo16, lo16b, reslo; // channel offsets of the residual ("lo") planes of out16 / out16b / res16 (0 = none)
  int num_m_tiles, num_n_tiles;
  // Grouped launch (HoVer-Net's three decoder branches in one grid): the M rows are `groups` consecutive blocks of
  // group_rows rows; group g uses weight rows [g*N, (g+1)*N), bias/scale2/shift2 + g*N and pre_scale/pre_shift + g*cin.
  // Tiles never straddle groups (each group has its own ceil(group_rows/128) row tiles).  a_grouped == 0: all groups read
  // the same A rows [0, group_rows) (a shared input tensor).
  int groups, group_rows, a_grouped, m_tiles_pg;
  // Fused pointwise head (HoVer-Net u0, hovernet.py:487-497): out[j] = b[j] + sum_c relu(acc[c]*scale[c] + shift[c]) * w[j][c] over the
  // 64 output channels of a BLOCK_N == 64 tile, j < nout <= 8, written as fp32 [rows, head_ld] per group; nothing else is stored.
  const float* head_w;      // [groups][8][64]
  const float* head_b;      // [groups][8]
  const float* head_scale;  // [groups][64]
  const float* head_shift;  // [groups][64]
  float* head_out[4];
  int head_nout[4], head_ld[4];
  int epi_mode;           // specialised epilogue row loop (0 = generic), see tiab_conv_gemm
  int skip;               // 1: the window does not span all grid rows -> tiles with no valid output row are skipped
};

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cvta_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbarrier_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(cvta_smem(bar)), "r"(count));
}
__device__ __forceinline__ void mbarrier_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(cvta_smem(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbarrier_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(cvta_smem(bar)) : "memory");
}
__device__ __forceinline__ void mbarrier_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "W_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra D_%=;\n\t"
      "bra W_%=;\n\t"
      "D_%=:\n\t}" ::"r"(cvta_smem(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          cvta_smem(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(cvta_smem(bar))
      : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(cvta_smem(bar))
               : "memory");
}
__device__ __forceinline__ void tcgen05_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100 "version 1"):
//   [0,14) start address >> 4 ; [16,30) leading byte offset >> 4 (= 1, unused for swizzled K-major) ;
//   [32,46) stride byte offset >> 4 (= 1024 B between 8-row groups) ; [46,48) version = 1 ;
//   [61,64) layout type = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// explicit shared-state-space accesses (32-bit addresses, immediate offsets; generic LD/ST cost extra address checks)
__device__ __forceinline__ float4 lds_f4(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ uint2 lds_u2(uint32_t a) {
  uint2 v;
  asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts_u2(uint32_t a, uint32_t x, uint32_t y) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(a), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void sts_u4(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_h2(uint32_t u) { return __half22float2(*reinterpret_cast<const __half2*>(&u)); }

// Phase 2 of the operand transform for one stage: 16 row passes of this thread's 4 channels, relu?(x*sc + sh) -> fp16 hi
// (+ lo) at the SWIZZLE_128B position.  CHECK adds the per-row range test (tiles at the ends of the tensor only); the two
// source formats and the checked variant are separate instantiations so that the hot path stays short (the three warp
// roles of this kernel share a 32 KB instruction cache).
template <int NSEG, bool F32, bool CHECK>
__device__ __forceinline__ void transform_store(const float4 (&xf)[BLOCK_M / 8], uint32_t dst, uint32_t lo_off, float4 sc, float4 sh,
                                                bool relu, int rsub, int r_lo, int r_hi) {
#pragma unroll
  for (int pass = 0; pass < BLOCK_M / 8; ++pass) {
    float v0, v1, v2, v3;
    if (F32) {
      v0 = xf[pass].x, v1 = xf[pass].y, v2 = xf[pass].z, v3 = xf[pass].w;
    } else {
      const float2 f0 = unpack_h2(__float_as_uint(xf[pass].x)), f1 = unpack_h2(__float_as_uint(xf[pass].y));
      const float2 g0 = unpack_h2(__float_as_uint(xf[pass].z)), g1 = unpack_h2(__float_as_uint(xf[pass].w));
      v0 = f0.x + g0.x, v1 = f0.y + g0.y, v2 = f1.x + g1.x, v3 = f1.y + g1.y;
    }
    v0 = fmaf(v0, sc.x, sh.x), v1 = fmaf(v1, sc.y, sh.y), v2 = fmaf(v2, sc.z, sh.z), v3 = fmaf(v3, sc.w, sh.w);
    if (relu) v0 = fmaxf(v0, 0.f), v1 = fmaxf(v1, 0.f), v2 = fmaxf(v2, 0.f), v3 = fmaxf(v3, 0.f);
    if (CHECK) {
      const int r = pass * 8 + rsub;
      if (r < r_lo || r >= r_hi) v0 = v1 = v2 = v3 = 0.f;
    }
    const uint32_t h01 = pack_h2(v0, v1), h23 = pack_h2(v2, v3);
    sts_u2(dst + pass * 1024, h01, h23);   // row r = pass*8 + rsub: 8-row groups are 1024 B apart
    if (NSEG == 3) {
      const float2 b01 = unpack_h2(h01), b23 = unpack_h2(h23);
      sts_u2(dst + lo_off + pass * 1024, pack_h2(v0 - b01.x, v1 - b01.y), pack_h2(v2 - b23.x, v3 - b23.y));
    }
  }
}

// Coalesced half of the epilogue for one 32-column chunk: lane (rsub, cq) handles rows i*4 + rsub (i = 0..7), columns
// 4*cq..4*cq+3 of the transposed staging tile.  RES: 0 none / 1 fp32 residual (prefetched in rres, or loaded in place when
// !PRE) / 2 fp16 hi(+lo) residual / -1 decide at run time; O32 / O16 / O16B select the outputs (GENERIC: all run-time flags).
// The common combinations are separate instantiations so that the executed path has no per-row feature branches.
template <bool GENERIC, int RES, bool O32, bool O16, bool O16B, bool LO, bool PRE>
__device__ __forceinline__ void epilogue_rows(const ConvParams& p, uint32_t stg_s, int rsub, int cq, uint32_t vmask, const int (&drow)[8],
                                              const int (&rrow)[8], const float4 (&rres)[8], int n, float4 os4, float4 bias4, float4 sc2,
                                              float4 sh2, float floor_) {
  const bool res32 = GENERIC ? p.res32 != nullptr : RES == 1;
  const bool res16 = GENERIC ? (p.res32 == nullptr && p.res16 != nullptr) : RES == 2;
  const bool o32 = GENERIC ? p.out32 != nullptr : O32;
  const bool o16 = GENERIC ? p.out16 != nullptr : O16;
  const bool o16b = GENERIC ? p.out16b != nullptr : O16B;
  const bool lo_a = GENERIC ? p.lo16 != 0 : LO, lo_b = GENERIC ? p.lo16b != 0 : LO;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (!(vmask >> i & 1u)) continue;
    const int rl7 = (i & 1) * 4 + rsub;   // (i*4 + rsub) & 7
    float4 x = lds_f4(stg_s + i * 512 + rsub * 128 + ((cq ^ rl7) << 4));
    x.x = fmaf(x.x, os4.x, bias4.x), x.y = fmaf(x.y, os4.y, bias4.y), x.z = fmaf(x.z, os4.z, bias4.z), x.w = fmaf(x.w, os4.w, bias4.w);
    if (res32) {
      const float4 rr = PRE ? rres[i] : *reinterpret_cast<const float4*>(p.res32 + (long long)rrow[i] * p.ldres + p.coffres + n);
      x.x += rr.x, x.y += rr.y, x.z += rr.z, x.w += rr.w;
    } else if (res16) {
      const __half* rp = p.res16 + (long long)rrow[i] * p.ldres + p.coffres + n;
      const uint2 q = *reinterpret_cast<const uint2*>(rp);
      const float2 f0 = unpack_h2(q.x), f1 = unpack_h2(q.y);
      x.x += f0.x, x.y += f0.y, x.z += f1.x, x.w += f1.y;
      if (p.reslo) {
        const uint2 ql = *reinterpret_cast<const uint2*>(rp + p.reslo);
        const float2 g0 = unpack_h2(ql.x), g1 = unpack_h2(ql.y);
        x.x += g0.x, x.y += g0.y, x.z += g1.x, x.w += g1.y;
      }
    }
    x.x = fmaxf(x.x, floor_), x.y = fmaxf(x.y, floor_), x.z = fmaxf(x.z, floor_), x.w = fmaxf(x.w, floor_);   // ReLU or no-op (-inf)
    if (o32) *reinterpret_cast<float4*>(p.out32 + (long long)drow[i] * p.ld32 + p.coff32 + n) = x;
    if (o16) {
      __half* op = p.out16 + (long long)(drow[i] & 0x0FFFFFFF) * p.ld16 + p.coff16 + (drow[i] >> 28) * p.s2d_c + n;
      const uint32_t h01 = pack_h2(x.x, x.y), h23 = pack_h2(x.z, x.w);
      *reinterpret_cast<uint2*>(op) = make_uint2(h01, h23);
      if (lo_a) {
        const float2 b01 = unpack_h2(h01), b23 = unpack_h2(h23);
        *reinterpret_cast<uint2*>(op + p.lo16) = make_uint2(pack_h2(x.x - b01.x, x.y - b01.y), pack_h2(x.z - b23.x, x.w - b23.y));
      }
    }
    if (o16b) {
      __half* op = p.out16b + (long long)drow[i] * p.ld16b + p.coff16b + n;
      const float a0 = fmaxf(fmaf(x.x, sc2.x, sh2.x), 0.f), a1 = fmaxf(fmaf(x.y, sc2.y, sh2.y), 0.f);
      const float a2 = fmaxf(fmaf(x.z, sc2.z, sh2.z), 0.f), a3 = fmaxf(fmaf(x.w, sc2.w, sh2.w), 0.f);
      const uint32_t h01 = pack_h2(a0, a1), h23 = pack_h2(a2, a3);
      *reinterpret_cast<uint2*>(op) = make_uint2(h01, h23);
      if (lo_b) {
        const float2 b01 = unpack_h2(h01), b23 = unpack_h2(h23);
        *reinterpret_cast<uint2*>(op + p.lo16b) = make_uint2(pack_h2(a0 - b01.x, a1 - b01.y), pack_h2(a2 - b23.x, a3 - b23.y));
      }
    }
  }
}

// True when at least one of the tile's 128 grid rows lies inside the valid y-window (every warp role evaluates this for
// the same tile sequence, so skipped tiles never touch a barrier).
// Row tile m_blk -> its group g, first global output row m0 and the end (exclusive) of the group's rows.
__device__ __forceinline__ void tile_rows(const ConvParams& p, int m_blk, int& g, int& m0, int& mend) {
  g = m_blk / p.m_tiles_pg;
  m0 = g * p.group_rows + (m_blk - g * p.m_tiles_pg) * BLOCK_M;
  mend = (g + 1) * p.group_rows;
}

__device__ __forceinline__ bool tile_active(const ConvParams& p, int m_blk) {
  if (!p.skip) return true;
  const int gp = p.Hin * p.Win;
  int g, r0, mend;
  tile_rows(p, m_blk, g, r0, mend);
  const int r1 = min(r0 + BLOCK_M, mend) - 1;
  const int i0 = r0 / gp, i1 = r1 / gp;
  const int ya = (r0 - i0 * gp) / p.Win, yb = (r1 - i1 * gp) / p.Win;
  if (i0 == i1) return !(yb < p.y0 || ya >= p.y1);
  if (i1 == i0 + 1) return !(ya >= p.y1 && yb < p.y0);
  return true;
}

// One pipeline stage holds everything ONE K-chunk of one tap needs:
//   NSEG == 1 : [A 16 KB][B]                      -> 4 UMMAs  (A*W)
//   NSEG == 3 : [A_hi 16 KB][A_lo 16 KB][W_hi][W_lo] -> 12 UMMAs (A_hi*W_hi + A_lo*W_hi + A_hi*W_lo), every operand loaded once
//   AMODE == 1: the A area is always 32 KB: TMA drops the RAW tile there (fp32 [128 x 64] or fp16 hi + lo planes), the
//               transform warps pull it into registers and overwrite it in place with the swizzled fp16 hi (+ lo) tiles.
template <int BLOCK_N, int NSEG, int AMODE>
struct ConvSmem {
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2;
  static constexpr int NPLANES = NSEG == 3 ? 2 : 1;
  static constexpr int A_AREA = (NSEG == 3 || AMODE) ? 2 * A_BYTES : A_BYTES;
  static constexpr int STAGE_BYTES = A_AREA + NPLANES * B_BYTES;
  static constexpr int STG_BYTES = (AMODE ? 4 : 8) * 4096;   // one 32x32 fp32 transpose tile per epilogue warp
  static constexpr int BUDGET = 227 * 1024 - 1024 - 384 - STG_BYTES;
  static constexpr int STAGES = (BUDGET / STAGE_BYTES) > MAX_STAGES ? MAX_STAGES : (BUDGET / STAGE_BYTES);
  static constexpr int STG_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int BAR_OFFSET = STG_OFFSET + STG_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + 384 + 1024;  // barriers + tmem ptr, + slack for 1024 B alignment
  // Narrow split-precision tiles (N <= 64) are bound by the shared-memory reads of the A operand (4 KB per UMMA whatever N
  // is), so W_hi and W_lo -- adjacent in the stage -- are fed as ONE 2N-row B tile: A_hi x [W_hi; W_lo] lands in 2N
  // accumulator columns, A_lo x W_hi adds to the first N, and the epilogue sums the two halves (2 UMMAs per K step, not 3).
  static constexpr bool CONCAT = NSEG == 3 && BLOCK_N <= 64;
  static constexpr int ACC_COLS = CONCAT ? 2 * BLOCK_N : BLOCK_N;
  static constexpr int TMEM_COLS = (2 * ACC_COLS) <= 32 ? 32 : (2 * ACC_COLS <= 64 ? 64 : (2 * ACC_COLS <= 128 ? 128 : (2 * ACC_COLS <= 256 ? 256 : 512)));
};

template <int BLOCK_N, int NSEG, int AMODE>
__global__ void __launch_bounds__(AMODE ? CONV_THREADS_T : CONV_THREADS, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
                 const __grid_constant__ CUtensorMap tmB, const ConvParams p) {
  using S = ConvSmem<BLOCK_N, NSEG, AMODE>;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + MAX_STAGES;
  uint64_t* tfull_bar = empty_bar + MAX_STAGES;   // [2]
  uint64_t* tempty_bar = tfull_bar + 2;           // [2]
  // (AMODE) "raw A tile landed", one set per transform group: the groups take alternate K-iterations, so with an odd
  // stage count a stage alternates between them -- on a shared barrier a group one ring ahead of the other would see the
  // previous phase's parity and run on a tile that has not arrived.  Per-group barriers are only ever waited on in order.
  uint64_t* raw_bar = tempty_bar + 2;             // [2][MAX_STAGES]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(raw_bar + 2 * MAX_STAGES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.num_m_tiles * p.num_n_tiles;
  const int kiters = p.taps * p.kchunks;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA2)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < S::STAGES; ++s) {
      mbarrier_init(&full_bar[s], AMODE ? 5 : 1);  // TMA expect_tx arrive (+ one arrive per transform warp)
      mbarrier_init(&empty_bar[s], 1);
      mbarrier_init(&raw_bar[s], 1);
      mbarrier_init(&raw_bar[MAX_STAGES + s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbarrier_init(&tfull_bar[b], 1);
      mbarrier_init(&tempty_bar[b], AMODE ? 4 : 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(cvta_smem(tmem_ptr_smem)),
                 "n"(S::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===================== TMA producer =====================
    int s = 0;
    uint32_t ph = 0;
    int iter = 0;   // K-iterations issued so far (parity = transform group that will take the stage)
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int n_blk = tile % p.num_n_tiles, m_blk = tile / p.num_n_tiles;
      if (!tile_active(p, m_blk)) continue;
      int g, m0, mend;
      tile_rows(p, m_blk, g, m0, mend);
      if (!p.a_grouped) m0 -= g * p.group_rows;        // shared input: every group reads rows [0, group_rows)
      const int n0 = g * p.N + n_blk * BLOCK_N;        // weight rows of this group
      if (lane == 0) {
        for (int t = 0; t < p.taps; ++t) {
          const int row = m0 + p.tap_off[t];
          const int cbase = p.tap_coff[t];
          // weights of tap t: NSEG == 3 -> [W_hi | W_lo], each kchunks*64 columns
          const int wbase = t * S::NPLANES * p.kchunks;
          for (int kc = 0; kc < p.kchunks; ++kc, ++iter) {
            mbarrier_wait(&empty_bar[s], ph ^ 1u);
            uint8_t* st = smem + s * S::STAGE_BYTES;
            if (AMODE) {
              uint64_t* rb = &raw_bar[(iter & 1) * MAX_STAGES + s];
              // raw tile -> A area (rows / channels outside the tensor are zero filled), signalled on its own barrier
              const bool two = !p.a_is_f32 && p.a_lo;
              mbarrier_expect_tx(rb, (p.a_is_f32 || two) ? 2 * S::A_BYTES : S::A_BYTES);
              tma_load_2d(st, &tmA, cbase + kc * BLOCK_K, row, rb);
              if (two) tma_load_2d(st + S::A_BYTES, &tmA2, cbase + kc * BLOCK_K, row, rb);
              mbarrier_expect_tx(&full_bar[s], S::NPLANES * S::B_BYTES);
            } else {
              mbarrier_expect_tx(&full_bar[s], S::STAGE_BYTES);
              tma_load_2d(st, &tmA, cbase + kc * BLOCK_K, row, &full_bar[s]);
              if (NSEG == 3) tma_load_2d(st + S::A_BYTES, &tmA2, cbase + kc * BLOCK_K, row, &full_bar[s]);
            }
            tma_load_2d(st + S::A_AREA, &tmB, (wbase + kc) * BLOCK_K, n0, &full_bar[s]);
            if (NSEG == 3) tma_load_2d(st + S::A_AREA + S::B_BYTES, &tmB, (wbase + p.kchunks + kc) * BLOCK_K, n0, &full_bar[s]);
            if (++s == S::STAGES) {
              s = 0;
              ph ^= 1u;
            }
          }
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      // instruction descriptor: D=f32 (bit 4), A=B=f16 (0), K-major both, N>>3 at [17,23), M>>4 at [24,29)
      const uint32_t idesc = (1u << 4) | ((uint32_t)(BLOCK_N >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
      const uint32_t idesc2 = (1u << 4) | ((uint32_t)((2 * BLOCK_N) >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);  // N = 2*BLOCK_N
      int s = 0;
      uint32_t ph = 0;
      int ab = 0;
      uint32_t aph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        if (!tile_active(p, tile / p.num_n_tiles)) continue;
        mbarrier_wait(&tempty_bar[ab], aph ^ 1u);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(ab * S::ACC_COLS);
        for (int k = 0; k < kiters; ++k) {
          mbarrier_wait(&full_bar[s], ph);
          tcgen05_fence_after();
          const uint32_t base = cvta_smem(smem + s * S::STAGE_BYTES);
          if (NSEG == 3) {
            const uint64_t ahi = make_smem_desc(base), alo = make_smem_desc(base + S::A_BYTES);
            const uint64_t bhi = make_smem_desc(base + S::A_AREA), blo = make_smem_desc(base + S::A_AREA + S::B_BYTES);
            if (S::CONCAT) {
#pragma unroll
              for (int kk = 0; kk < BLOCK_K / UMMA_K; ++kk) {  // +2 in (addr >> 4) units = 32 bytes = 16 fp16 along K
                tcgen05_mma_f16(tmem_d, ahi + (uint64_t)(kk * 2), bhi + (uint64_t)(kk * 2), idesc2, (k > 0 || kk > 0) ? 1u : 0u);
                tcgen05_mma_f16(tmem_d, alo + (uint64_t)(kk * 2), bhi + (uint64_t)(kk * 2), idesc, 1u);
              }
            } else {
#pragma unroll
              for (int kk = 0; kk < BLOCK_K / UMMA_K; ++kk)
                tcgen05_mma_f16(tmem_d, ahi + (uint64_t)(kk * 2), bhi + (uint64_t)(kk * 2), idesc, (k > 0 || kk > 0) ? 1u : 0u);
#pragma unroll
              for (int kk = 0; kk < BLOCK_K / UMMA_K; ++kk)
                tcgen05_mma_f16(tmem_d, alo + (uint64_t)(kk * 2), bhi + (uint64_t)(kk * 2), idesc, 1u);
#pragma unroll
              for (int kk = 0; kk < BLOCK_K / UMMA_K; ++kk)
                tcgen05_mma_f16(tmem_d, ahi + (uint64_t)(kk * 2), blo + (uint64_t)(kk * 2), idesc, 1u);
            }
          } else {
            const uint64_t adesc = make_smem_desc(base), bdesc = make_smem_desc(base + S::A_AREA);
#pragma unroll
            for (int kk = 0; kk < BLOCK_K / UMMA_K; ++kk)
              tcgen05_mma_f16(tmem_d, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc, (k > 0 || kk > 0) ? 1u : 0u);
          }
          tcgen05_commit(&empty_bar[s]);  // frees the smem stage when these MMAs have consumed it
          if (++s == S::STAGES) {
            s = 0;
            ph ^= 1u;
          }
        }
        tcgen05_commit(&tfull_bar[ab]);  // accumulator ready for the epilogue
        if (++ab == 2) {
          ab = 0;
          aph ^= 1u;
        }
      }
    }
  } else if (warp >= 4 && (!AMODE || warp < 8)) {
    // ===================== epilogue: TMEM -> registers -> (smem transpose) -> global =====================
    // tcgen05.ld hands every lane one tile ROW; written like that, each global access of a warp touches 32 different
    // cache lines and the L1 tag stage (one line per cycle) becomes the bound of the memory-bound 1x1 convolutions.  Each warp
    // therefore transposes its 32x32 fp32 chunk through a private 4 KB swizzled staging tile so that 8 lanes cover 128
    // contiguous bytes of one output row (4 lines per request instead of 32).
    const int quad = warp & 3;           // TMEM lane quadrant this warp may access
    const int half = AMODE ? 0 : (warp - 4) >> 2;  // (8-warp epilogue) the two warps of a quadrant take alternate 32-column chunks
    const uint32_t stg_s = cvta_smem(smem + S::STG_OFFSET + (warp - 4) * 4096);
    const float floor_ = p.relu ? 0.f : -__int_as_float(0x7f800000);   // max(x, -inf) == x
    const int rsub = lane >> 3;          // coalesced phase: row i*4 + rsub of the chunk, ...
    const int cq = lane & 7;             // ... 16-byte column group cq (columns 4*cq .. 4*cq+3)
    int ab = 0;
    uint32_t aph = 0;
    const int grid_px = p.Hin * p.Win;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int n_blk = tile % p.num_n_tiles, m_blk = tile / p.num_n_tiles;
      if (!tile_active(p, m_blk)) continue;
      const int n0 = n_blk * BLOCK_N;
      int g, m0, mend;
      tile_rows(p, m_blk, g, m0, mend);
      const int gn = g * p.N;   // offset of this group's per-channel vectors
      // geometry of the row this lane owns in TMEM, then redistributed to the lanes that will store it
      int drow_l = 0, rrow_l = 0, valid_l = 0;
      {
        const int row = m0 + quad * 32 + lane;
        const int img = row / grid_px;
        const int rem = row - img * grid_px;
        const int y = rem / p.Win;
        const int x = rem - y * p.Win;
        valid_l = row < mend && y >= p.y0 && y < p.y1 && x >= p.x0 && x < p.x1;
        if (p.s2d_c) {   // space-to-depth store: cell (yy/2, xx/2), phase (yy%2, xx%2) carried in the top bits of the row index
          const int yy = y + p.dy, xx = x + p.dx;
          drow_l = ((img * p.Hd + (yy >> 1)) * p.Wd + (xx >> 1)) | ((((yy & 1) << 1) | (xx & 1)) << 28);
        } else {
          drow_l = (img * p.Hd + (y + p.dy)) * p.Wd + (x + p.dx);
        }
        rrow_l = (img * p.Hr + (y + p.ry)) * p.Wr + (x + p.rx);
      }
      int drow[8], rrow[8];
      uint32_t vmask = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int src = i * 4 + rsub;
        drow[i] = __shfl_sync(0xffffffffu, drow_l, src);
        rrow[i] = __shfl_sync(0xffffffffu, rrow_l, src);
        if (__shfl_sync(0xffffffffu, valid_l, src)) vmask |= 1u << i;
      }
      // fp32 residual: keep the NEXT chunk's loads in flight while the current one is combined and stored, and issue the
      // first chunk before waiting for the MMAs
      constexpr int CSTEP = AMODE ? 32 : 64;
      const float* rbase = p.res32 ? p.res32 + p.coffres + n0 + 4 * cq : nullptr;
      float4 rcur[8], rnext[8];
      if (!AMODE && rbase) {   // (the 512-thread transform variant has no registers to spare: it loads in place)
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (vmask >> i & 1u) rcur[i] = *reinterpret_cast<const float4*>(rbase + (long long)rrow[i] * p.ldres + half * 32);
      }
      mbarrier_wait(&tfull_bar[ab], aph);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(ab * S::ACC_COLS);
      bool released = false;
#pragma unroll 1
      for (int c0 = half * 32; c0 < BLOCK_N; c0 += CSTEP) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(taddr + (uint32_t)c0, r);
        if (!AMODE && rbase && c0 + CSTEP < BLOCK_N) {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (vmask >> i & 1u) rnext[i] = *reinterpret_cast<const float4*>(rbase + (long long)rrow[i] * p.ldres + c0 + CSTEP);
        }
        if (S::CONCAT) {   // second accumulator half (A_hi x W_lo)
          uint32_t r2[32];
          tmem_ld_32x32b_x32(taddr + (uint32_t)(BLOCK_N + c0), r2);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(r2[j]));
        } else {
          tmem_ld_wait();
        }
        if (c0 + CSTEP >= BLOCK_N) {   // last chunk of this warp is in registers: hand the accumulator back to the MMA warp now
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbarrier_arrive(&tempty_bar[ab]);
          released = true;
        }
        if (BLOCK_N == 64 && !AMODE && p.head_w != nullptr) {
          // ---- fused head: this warp holds channels [c0, c0+32) of its 32 rows; the quadrant's other warp holds the rest ----
          const int nout = p.head_nout[g];
          const float* hw = p.head_w + g * 8 * 64 + c0;
          const float* hs = p.head_scale + g * 64 + c0;
          const float* ht = p.head_shift + g * 64 + c0;
          float part[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) part[j] = 0.f;
#pragma unroll
          for (int c = 0; c < 32; c += 4) {
            float4 s4 = __ldg(reinterpret_cast<const float4*>(hs + c));
            const float4 t4 = __ldg(reinterpret_cast<const float4*>(ht + c));
            if (p.oscale) {   // fold the weight pre-scaling into the head's BN scale
              const float4 o4 = __ldg(reinterpret_cast<const float4*>(p.oscale + gn + c0 + c));
              s4.x *= o4.x, s4.y *= o4.y, s4.z *= o4.z, s4.w *= o4.w;
            }
            const float a0 = fmaxf(fmaf(__uint_as_float(r[c]), s4.x, t4.x), 0.f), a1 = fmaxf(fmaf(__uint_as_float(r[c + 1]), s4.y, t4.y), 0.f);
            const float a2 = fmaxf(fmaf(__uint_as_float(r[c + 2]), s4.z, t4.z), 0.f), a3 = fmaxf(fmaf(__uint_as_float(r[c + 3]), s4.w, t4.w), 0.f);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              if (j < nout) {
                const float4 w4 = __ldg(reinterpret_cast<const float4*>(hw + j * 64 + c));
                part[j] = fmaf(a0, w4.x, fmaf(a1, w4.y, fmaf(a2, w4.z, fmaf(a3, w4.w, part[j]))));
              }
            }
          }
          // combine the two column halves through the upper warp's staging tile (32 bytes per row)
          const uint32_t xch = cvta_smem(smem + S::STG_OFFSET + (4 + quad) * 4096) + lane * 32;
          if (half == 1) {
            sts_u4(xch, __float_as_uint(part[0]), __float_as_uint(part[1]), __float_as_uint(part[2]), __float_as_uint(part[3]));
            sts_u4(xch + 16, __float_as_uint(part[4]), __float_as_uint(part[5]), __float_as_uint(part[6]), __float_as_uint(part[7]));
          }
          asm volatile("bar.sync %0, 64;" ::"r"(quad + 1) : "memory");
          if (half == 0) {
            const float4 q0 = lds_f4(xch), q1 = lds_f4(xch + 16);
            const float other[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
            if (valid_l) {
              const long long lrow = (long long)drow_l - (long long)g * (p.group_rows / grid_px) * p.Hd * p.Wd;   // row inside this group's output
              float* op = p.head_out[g] + lrow * p.head_ld[g];
              const float* hb = p.head_b + g * 8;
#pragma unroll
              for (int j = 0; j < 8; ++j)
                if (j < nout) op[j] = part[j] + other[j] + __ldg(hb + j);
            }
          }
          asm volatile("bar.sync %0, 64;" ::"r"(quad + 1) : "memory");   // the exchange slots are rewritten by the next tile
          continue;
        }
        // transpose through the staging tile: 16-byte group g of row `lane` lives at lane*128 + ((g ^ (lane & 7)) << 4)
#pragma unroll
        for (int g = 0; g < 8; ++g) sts_u4(stg_s + lane * 128 + ((g ^ (lane & 7)) << 4), r[4 * g], r[4 * g + 1], r[4 * g + 2], r[4 * g + 3]);
        __syncwarp();
        const int n = n0 + c0 + 4 * cq;   // first of this lane's 4 output channels
        float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f), sc2 = bias4, sh2 = bias4;
        if (p.bias) bias4 = __ldg(reinterpret_cast<const float4*>(p.bias + gn + n));
        float4 os4 = make_float4(1.f, 1.f, 1.f, 1.f);
        if (p.oscale) os4 = __ldg(reinterpret_cast<const float4*>(p.oscale + gn + n));
        if (p.out16b) {
          sc2 = __ldg(reinterpret_cast<const float4*>(p.scale2 + gn + n));
          sh2 = __ldg(reinterpret_cast<const float4*>(p.shift2 + gn + n));
        }
        constexpr bool PRE = !AMODE;
        switch (p.epi_mode) {
          case 1: epilogue_rows<false, 0, false, true, false, true, PRE>(p, stg_s, rsub, cq, vmask, drow, rrow, rcur, n, os4, bias4, sc2, sh2, floor_); break;
          case 2: epilogue_rows<false, 0, false, true, false, false, PRE>(p, stg_s, rsub, cq, vmask, drow, rrow, rcur, n, os4, bias4, sc2, sh2, floor_); break;
          case 3: epilogue_rows<false, 0, true, false, false, false, PRE>(p, stg_s, rsub, cq, vmask, drow, rrow, rcur, n, os4, bias4, sc2, sh2, floor_); break;
          case 4: epilogue_rows<false, 1, true, false, false, false, PRE>(p, stg_s, rsub, cq, vmask, drow, rrow, rcur, n, os4, bias4, sc2, sh2, floor_); break;
          case 5: epilogue_rows<false, 1, false, false, true, true, PRE>(p, stg_s, rsub, cq, vmask, drow, rrow, rcur, n, os4, bias4, sc2, sh2, floor_); break;
          default: epilogue_rows<true, -1, false, false, false, false, PRE>(p, stg_s, rsub, cq, vmask, drow, rrow, rcur, n, os4, bias4, sc2, sh2, floor_); break;
        }
        if (!AMODE) {
#pragma unroll
          for (int i = 0; i < 8; ++i) rcur[i] = rnext[i];
        }
        __syncwarp();   // the staging tile is rewritten by the next chunk
      }
      if (!released) {   // (warps without a chunk: BLOCK_N == 32, second half)
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbarrier_arrive(&tempty_bar[ab]);
      }
      if (++ab == 2) {
        ab = 0;
        aph ^= 1u;
      }
    }
  }

  if (AMODE && warp >= 8) {
    // ===================== A-operand transform warps: raw tensor -> BN/ReLU -> hi/lo swizzled smem tiles =====================
    const int tw = (threadIdx.x - 256) & 127;   // 0..127 inside the group
    const int grp = (threadIdx.x - 256) >> 7;   // the two groups take alternate K-iterations so that two stages are in flight
    int iter = 0;
    const int rsub = tw >> 4;           // row within an 8-row pass (16 threads cover one 64-channel row chunk)
    const int c4 = (tw & 15) * 4;       // channel offset inside the 64-channel chunk
    const uint32_t smem_s = cvta_smem(smem);
    uint32_t rawph = 0;                 // bit s: parity of this group's next wait on its raw barrier of stage s
    int s = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile / p.num_n_tiles;
      if (!tile_active(p, m_blk)) continue;
      int g, m0, mend;
      tile_rows(p, m_blk, g, m0, mend);
      if (!p.a_grouped) m0 -= g * p.group_rows;
      const int gc = g * p.a_cin;   // this group's pre-activation vectors
      for (int t = 0; t < p.taps; ++t) {
        const long long row0 = (long long)m0 + p.tap_off[t];
        for (int kc = 0; kc < p.kchunks; ++kc, ++iter) {
          if ((iter & 1) != grp) {  // the other group's iteration: just keep the stage/phase bookkeeping in step
            if (++s == S::STAGES) {
              s = 0;
              ph ^= 1u;
            }
            continue;
          }
          const int ch = kc * BLOCK_K + c4;
          const bool chok = ch < p.a_cin;
          // channels beyond cin (K padding; TMA zero-fills them) get scale = shift = 0 -> operand 0, without a divergent path
          float4 sc = chok ? make_float4(1.f, 1.f, 1.f, 1.f) : make_float4(0.f, 0.f, 0.f, 0.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
          if (p.pre_scale && chok) {
            sc = __ldg(reinterpret_cast<const float4*>(p.pre_scale + gc + ch));
            sh = __ldg(reinterpret_cast<const float4*>(p.pre_shift + gc + ch));
          }
          const uint32_t st = smem_s + (uint32_t)(s * S::STAGE_BYTES);
          // phase 1: pull this thread's share of the raw tile (TMA-staged) into registers; phase 2 (after the whole group
          // has read): transform and overwrite the same bytes with the swizzled hi / lo operand tiles
          constexpr int NP = BLOCK_M / 8;
          float4 xf[NP];
          mbarrier_wait(&raw_bar[grp * MAX_STAGES + s], (rawph >> s) & 1u);
          rawph ^= 1u << s;
          if (p.a_is_f32) {
#pragma unroll
            for (int pass = 0; pass < NP; ++pass) xf[pass] = lds_f4(st + rsub * (BLOCK_K * 4) + c4 * 4 + pass * (8 * BLOCK_K * 4));
          } else {
#pragma unroll
            for (int pass = 0; pass < NP; ++pass) {
              const uint2 q = lds_u2(st + rsub * (BLOCK_K * 2) + c4 * 2 + pass * (8 * BLOCK_K * 2));
              uint2 ql = make_uint2(0u, 0u);
              if (p.a_lo) ql = lds_u2(st + S::A_BYTES + rsub * (BLOCK_K * 2) + c4 * 2 + pass * (8 * BLOCK_K * 2));
              xf[pass] = make_float4(__uint_as_float(q.x), __uint_as_float(q.y), __uint_as_float(ql.x), __uint_as_float(ql.y));
            }
          }
          asm volatile("bar.sync %0, 128;" ::"r"(grp + 1) : "memory");
          // rows [r_lo, r_hi) of this tile exist in the source tensor (all of them except at its two ends)
          const long long lo64 = -row0, hi64 = p.a_rows - row0;
          const int r_lo = lo64 > 0 ? (int)(lo64 < BLOCK_M ? lo64 : BLOCK_M) : 0;
          const int r_hi = hi64 >= BLOCK_M ? BLOCK_M : (hi64 > 0 ? (int)hi64 : 0);
          // SWIZZLE_128B: 16-byte chunk index XOR (row % 8) inside each 128-byte row
          const uint32_t dst = st + (uint32_t)(rsub * 128 + (((c4 >> 3) ^ rsub) << 4) + (c4 & 7) * 2);
          const bool full = r_lo == 0 && r_hi == BLOCK_M;
          if (p.a_is_f32) {
            if (full) transform_store<NSEG, true, false>(xf, dst, S::A_BYTES, sc, sh, p.pre_relu != 0, rsub, r_lo, r_hi);
            else transform_store<NSEG, true, true>(xf, dst, S::A_BYTES, sc, sh, p.pre_relu != 0, rsub, r_lo, r_hi);
          } else {
            if (full) transform_store<NSEG, false, false>(xf, dst, S::A_BYTES, sc, sh, p.pre_relu != 0, rsub, r_lo, r_hi);
            else transform_store<NSEG, false, true>(xf, dst, S::A_BYTES, sc, sh, p.pre_relu != 0, rsub, r_lo, r_hi);
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the tensor core (async proxy)
          __syncwarp();
          if (lane == 0) mbarrier_arrive(&full_bar[s]);
          if (++s == S::STAGES) {
            s = 0;
            ph ^= 1u;
          }
        }
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(S::TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// auxiliary kernels (memory-bound glue between the GEMMs).
// Activations may carry a second fp16 plane ("lo" = x - fp16(x)) at a fixed channel offset so that
// hi + lo represents the value to ~22 bits (split-precision mode); an offset of 0 means "no lo plane".
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void load8(const __half* p, int lo, float (&v)[8]) {
  uint4 q = *reinterpret_cast<const uint4*>(p);
  const __half2* h2 = reinterpret_cast<const __half2*>(&q);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    float2 f = __half22float2(h2[e]);
    v[2 * e] = f.x, v[2 * e + 1] = f.y;
  }
  if (lo) {
    uint4 ql = *reinterpret_cast<const uint4*>(p + lo);
    const __half2* l2 = reinterpret_cast<const __half2*>(&ql);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float2 f = __half22float2(l2[e]);
      v[2 * e] += f.x, v[2 * e + 1] += f.y;
    }
  }
}
__device__ __forceinline__ void store8(__half* p, int lo, const float (&v)[8]) {
  uint4 q, ql;
  __half2* h2 = reinterpret_cast<__half2*>(&q);
  __half2* l2 = reinterpret_cast<__half2*>(&ql);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    h2[e] = __floats2half2_rn(v[2 * e], v[2 * e + 1]);
    float2 hf = __half22float2(h2[e]);
    l2[e] = __floats2half2_rn(v[2 * e] - hf.x, v[2 * e + 1] - hf.y);
  }
  *reinterpret_cast<uint4*>(p) = q;
  if (lo) *reinterpret_cast<uint4*>(p + lo) = ql;
}

// Explicit im2col for strided convs / 3-channel stems: out[m, (kh*KW + kw)*C + c] (row pitch ldout, K zero padded to
// Kpad, lo plane at +out_lo) from an NHWC source that is fp16 (+lo plane) or uint8 (scaled by u8_scale).
// flattened-index decomposition of the memory-bound glue kernels: a 64-bit division costs ~100 instructions on the GPU, and these
// kernels move 16-48 bytes per thread — with 64-bit div/mod chains they were instruction-bound (upsample + skip add at ~55 % of the
// HBM roofline in r2h).  Indices below 2^32 (every launch of the three networks) take the 32-bit path.
__device__ __forceinline__ long long fast_divmod(long long i, int d, int& rem) {
  if ((unsigned long long)i <= 0xffffffffull) {
    const unsigned ui = (unsigned)i, q = ui / (unsigned)d;
    rem = (int)(ui - q * (unsigned)d);
    return (long long)q;
  }
  const long long q = i / d;
  rem = (int)(i - q * d);
  return q;
}

__global__ void __launch_bounds__(256) im2col_kernel(const void* __restrict__ src, int src_is_u8, float u8_scale, int src_lo,
                                                     int Nimg, int H, int W, int C, int ldsrc, int KH, int KW, int stride,
                                                     int pad_t, int pad_l, int Ho, int Wo, __half* __restrict__ out,
                                                     int ldout, int Kpad, int out_lo) {
  const long long total = (long long)Nimg * Ho * Wo * Kpad;
  const int K = KH * KW * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % Kpad);
    const long long m = i / Kpad;
    float v = 0.f;
    if (k < K) {
      const int c = k % C, t = k / C, kw = t % KW, kh = t / KW;
      const int xo = (int)(m % Wo);
      const int yo = (int)((m / Wo) % Ho);
      const int img = (int)(m / ((long long)Wo * Ho));
      const int y = yo * stride - pad_t + kh, x = xo * stride - pad_l + kw;
      if (y >= 0 && y < H && x >= 0 && x < W) {
        const long long off = ((long long)(img * H + y) * W + x) * ldsrc + c;
        if (src_is_u8) {
          v = (float)reinterpret_cast<const uint8_t*>(src)[off] * u8_scale;
        } else {
          const __half* sp = reinterpret_cast<const __half*>(src);
          v = __half2float(sp[off]);
          if (src_lo) v += __half2float(sp[off + src_lo]);
        }
      }
    }
    const __half hi = __float2half_rn(v);
    out[m * ldout + k] = hi;
    if (out_lo) out[m * ldout + out_lo + k] = __float2half_rn(v - __half2float(hi));
  }
}

// 3-channel uint8 stem (7x7 etc.): one thread per 8 consecutive K elements of an output pixel -> one 16-byte store per plane.
// The k -> (kh, kw, c) decomposition comes from a shared-memory table (no runtime divisions in the inner loop):
// entry = byte offset (kh*W + kw)*3 + c | kh << 24 | kw << 28, or 0xFFFFFFFF for the zero padding of K.
__global__ void __launch_bounds__(256) im2col_u8c3_kernel(const uint8_t* __restrict__ src, float u8_scale, int Nimg, int H, int W, int KH,
                                                          int KW, int stride, int pad_t, int pad_l, int Ho, int Wo, __half* __restrict__ out,
                                                          int ldout, int Kpad, int out_lo) {
  extern __shared__ uint32_t ktab[];
  const int K = KH * KW * 3;
  for (int k = threadIdx.x; k < Kpad; k += blockDim.x) {
    uint32_t e = 0xFFFFFFFFu;
    if (k < K) {
      const int c = k % 3, t = k / 3, kw = t % KW, kh = t / KW;
      e = (uint32_t)((kh * W + kw) * 3 + c) | ((uint32_t)kh << 24) | ((uint32_t)kw << 28);
    }
    ktab[k] = e;
  }
  __syncthreads();
  const int k8n = Kpad >> 3;
  const long long total = (long long)Nimg * Ho * Wo * k8n;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int g, xo, yo;
    const long long m = fast_divmod(i, k8n, g);
    const long long my = fast_divmod(m, Wo, xo);
    const int img = (int)fast_divmod(my, Ho, yo);
    const int ybase = yo * stride - pad_t, xbase = xo * stride - pad_l;
    const uint8_t* sp = src + ((long long)img * H * W + (long long)ybase * W + xbase) * 3;
    const bool interior = ybase >= 0 && xbase >= 0 && ybase + KH <= H && xbase + KW <= W;
    const uint4 e0 = *reinterpret_cast<const uint4*>(ktab + g * 8), e1 = *reinterpret_cast<const uint4*>(ktab + g * 8 + 4);
    const uint32_t ee[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float val = 0.f;
      const uint32_t t = ee[e];
      if (t != 0xFFFFFFFFu) {
        bool ok = interior;
        if (!interior) {
          const int y = ybase + (int)((t >> 24) & 15u), x = xbase + (int)(t >> 28);
          ok = y >= 0 && y < H && x >= 0 && x < W;
        }
        if (ok) val = (float)sp[t & 0xFFFFFFu] * u8_scale;
      }
      v[e] = val;
    }
    store8(out + m * ldout + g * 8, out_lo, v);
  }
}

// vectorised variant: fp16 source with C % 8 == 0 -> one 16-byte copy per 8 consecutive K elements (same tap)
__global__ void __launch_bounds__(256) im2col_vec8_kernel(const __half* __restrict__ src, int src_lo, int Nimg, int H, int W, int C,
                                                          int ldsrc, int KH, int KW, int stride, int pad_t, int pad_l, int Ho, int Wo,
                                                          __half* __restrict__ out, int ldout, int Kpad, int out_lo) {
  const int k8n = Kpad >> 3;
  const long long total = (long long)Nimg * Ho * Wo * k8n;
  const int K = KH * KW * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int k;
    const long long m = fast_divmod(i, k8n, k);
    k *= 8;
    uint4 hi = make_uint4(0u, 0u, 0u, 0u), lo = hi;
    if (k < K) {
      const int c = k % C, t = k / C, kw = t % KW, kh = t / KW;
      int xo, yo;
      const int img = (int)fast_divmod(fast_divmod(m, Wo, xo), Ho, yo);
      const int y = yo * stride - pad_t + kh, x = xo * stride - pad_l + kw;
      if (y >= 0 && y < H && x >= 0 && x < W) {
        const __half* sp = src + ((long long)(img * H + y) * W + x) * ldsrc + c;
        hi = *reinterpret_cast<const uint4*>(sp);
        if (src_lo) lo = *reinterpret_cast<const uint4*>(sp + src_lo);
      }
    }
    *reinterpret_cast<uint4*>(out + m * ldout + k) = hi;
    if (out_lo) *reinterpret_cast<uint4*>(out + m * ldout + out_lo + k) = lo;
  }
}

// y = relu?(x * scale[c] + shift[c]) over C channels of `rows` rows (fp32 or fp16(+lo) source) -> fp16(+lo).
// Per-unit pre-activation of HoVer-Net's dense blocks (hovernet.py:84-86).
__global__ void __launch_bounds__(256) bn_relu_kernel(const void* __restrict__ src, int src_is_f32, int ldsrc, int coffsrc,
                                                      int src_lo, const float* __restrict__ scale,
                                                      const float* __restrict__ shift, int relu, long long rows, int C,
                                                      __half* __restrict__ dst, int lddst, int coffdst, int dst_lo) {
  const int c8 = C >> 3;
  const long long total = rows * c8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int cg;
    const long long r = fast_divmod(i, c8, cg);
    cg *= 8;
    float v[8];
    if (src_is_f32) {
      const float* s = reinterpret_cast<const float*>(src) + r * ldsrc + coffsrc + cg;
      float4 a = *reinterpret_cast<const float4*>(s), b = *reinterpret_cast<const float4*>(s + 4);
      v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = b.x, v[5] = b.y, v[6] = b.z, v[7] = b.w;
    } else {
      load8(reinterpret_cast<const __half*>(src) + r * ldsrc + coffsrc + cg, src_lo, v);
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      if (scale) v[e] = fmaf(v[e], scale[cg + e], shift[cg + e]);
      if (relu) v[e] = fmaxf(v[e], 0.f);
    }
    store8(dst + r * lddst + coffdst + cg, dst_lo, v);
  }
}

// dst[n, y+dy, x+dx, :] = relu?(bn?(up2x(a window)[n, y, x, :] + b[n, y+by, x+bx, :]))  (nearest 2x, NHWC fp16(+lo))
// hovernet.py:440-451 `upsample2x(d[-1]) + d[-2]`, unet.py:405-411 (+ the decoder's BN-ReLU pre-activation, :211-224).
__global__ void __launch_bounds__(256)
upsample_add_kernel(const __half* __restrict__ a, int Ha, int Wa, int Hag, int Wag, int ay, int ax, int lda, int a_lo,
                    const __half* __restrict__ b, int Hb, int Wb, int by, int bx, int ldb, int b_lo, int Nimg, int C,
                    const float* __restrict__ scale, const float* __restrict__ shift, int relu, __half* __restrict__ dst,
                    int Hd, int Wd, int dy, int dx, int ldd, int dst_lo) {
  const int Ho = 2 * Ha, Wo = 2 * Wa, c8 = C >> 3;
  const long long total = (long long)Nimg * Ho * Wo * c8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int cg, x, y;
    long long r = fast_divmod(i, c8, cg);
    cg *= 8;
    r = fast_divmod(r, Wo, x);
    const int img = (int)fast_divmod(r, Ho, y);
    float v[8];
    load8(a + ((long long)(img * Hag + (y >> 1) + ay) * Wag + (x >> 1) + ax) * lda + cg, a_lo, v);
    if (b) {
      float w[8];
      load8(b + ((long long)(img * Hb + y + by) * Wb + (x + bx)) * ldb + cg, b_lo, w);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] += w[e];
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      if (scale) v[e] = fmaf(v[e], scale[cg + e], shift[cg + e]);
      if (relu) v[e] = fmaxf(v[e], 0.f);
    }
    store8(dst + ((long long)(img * Hd + y + dy) * Wd + (x + dx)) * ldd + cg, dst_lo, v);
  }
}

// 3x3 stride-2 pad-1 max pool, NHWC fp16(+lo) (torchvision ResNet stem), destination may be a padded window.
__global__ void __launch_bounds__(256) maxpool3x3s2_kernel(const __half* __restrict__ src, int Nimg, int H, int W, int C, int lds,
                                                           int src_lo, __half* __restrict__ dst, int Ho, int Wo, int Hd, int Wd,
                                                           int dy, int dx, int ldd, int dst_lo) {
  const int c8 = C >> 3;
  const long long total = (long long)Nimg * Ho * Wo * c8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int cg, xo, yo;
    long long r = fast_divmod(i, c8, cg);
    cg *= 8;
    r = fast_divmod(r, Wo, xo);
    const int img = (int)fast_divmod(r, Ho, yo);
    float m[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) m[e] = -3.0e38f;
    for (int ky = 0; ky < 3; ++ky)
      for (int kx = 0; kx < 3; ++kx) {
        const int y = 2 * yo - 1 + ky, x = 2 * xo - 1 + kx;
        if (y < 0 || y >= H || x < 0 || x >= W) continue;
        float v[8];
        load8(src + ((long long)(img * H + y) * W + x) * lds + cg, src_lo, v);
#pragma unroll
        for (int e = 0; e < 8; ++e) m[e] = fmaxf(m[e], v[e]);
      }
    store8(dst + ((long long)(img * Hd + yo + dy) * Wd + xo + dx) * ldd + cg, dst_lo, m);
  }
}

// Small-N pointwise head: out[p, j] = sum_c relu?(x[p,c]*scale[c]+shift[c]) * Wt[j, c] + b[j], j < nout <= 16, fp32 out.
// HoVer-Net u0 (hovernet.py:487-497: BN, ReLU, 1x1 conv with bias) and the UNet clf (unet.py:353).
__global__ void __launch_bounds__(128) pointwise_head_kernel(const __half* __restrict__ x, long long rows, int C, int ldx, int x_lo,
                                                             const float* __restrict__ scale, const float* __restrict__ shift,
                                                             int relu, const float* __restrict__ Wt, const float* __restrict__ b,
                                                             int nout, float* __restrict__ out, int ldout, int coffout) {
  extern __shared__ float sh[];  // Wt[nout*C] | scale[C] | shift[C]
  float* sw = sh;
  float* ss = sh + nout * C;
  float* st = ss + C;
  for (int i = threadIdx.x; i < nout * C; i += blockDim.x) sw[i] = Wt[i];
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    ss[i] = scale ? scale[i] : 1.f;
    st[i] = shift ? shift[i] : 0.f;
  }
  __syncthreads();
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (long long)gridDim.x * blockDim.x) {
    float acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = 0.f;
    const __half* xp = x + r * ldx;
    for (int c = 0; c < C; c += 8) {
      float v[8];
      load8(xp + c, x_lo, v);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float a = fmaf(v[e], ss[c + e], st[c + e]);
        if (relu) a = fmaxf(a, 0.f);
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (j < nout) acc[j] = fmaf(a, sw[j * C + c + e], acc[j]);
      }
    }
    for (int j = 0; j < nout; ++j) out[r * ldout + coffout + j] = acc[j] + (b ? b[j] : 0.f);
  }
}

// Global-average-pool + Linear + softmax (vanilla.py:312-316): x NHWC fp16(+lo) [N, HW, ldx] -> probs [N, ncls] fp32.
__global__ void __launch_bounds__(256) gap_fc_softmax_kernel(const __half* __restrict__ x, int HW, int C, int ldx, int x_lo,
                                                             const float* __restrict__ Wt, const float* __restrict__ b,
                                                             int ncls, float* __restrict__ feat_out, float* __restrict__ probs) {
  extern __shared__ float sh[];  // feat[C] | logits[ncls]
  float* feat = sh;
  float* logits = sh + C;
  const int img = blockIdx.x;
  const __half* xp = x + (long long)img * HW * ldx;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int i = 0; i < HW; ++i) {
      s += __half2float(xp[(long long)i * ldx + c]);
      if (x_lo) s += __half2float(xp[(long long)i * ldx + x_lo + c]);
    }
    feat[c] = s / (float)HW;
    if (feat_out) feat_out[(long long)img * C + c] = feat[c];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int j = warp; j < ncls; j += blockDim.x >> 5) {
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s = fmaf(feat[c], Wt[j * C + c], s);
    for (int d = 16; d; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if (lane == 0) logits[j] = s + b[j];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = -1e30f;
    for (int j = 0; j < ncls; ++j) m = fmaxf(m, logits[j]);
    float z = 0.f;
    for (int j = 0; j < ncls; ++j) z += expf(logits[j] - m);
    for (int j = 0; j < ncls; ++j) probs[(long long)img * ncls + j] = expf(logits[j] - m) / z;
  }
}

// HoVer-Net head activations (hovernet.py:897-909): np = softmax(np_logits)[..., 1], hv raw, tp = argmax as float.
__global__ void __launch_bounds__(256) hovernet_heads_kernel(const float* __restrict__ npl, const float* __restrict__ hvl,
                                                             const float* __restrict__ tpl, int nt, long long rows,
                                                             float* __restrict__ np_out, float* __restrict__ hv_out,
                                                             float* __restrict__ tp_out) {
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (long long)gridDim.x * blockDim.x) {
    const float l0 = npl[2 * r], l1 = npl[2 * r + 1];
    const float m = fmaxf(l0, l1);
    const float e0 = expf(l0 - m), e1 = expf(l1 - m);
    np_out[r] = e1 / (e0 + e1);
    hv_out[2 * r] = hvl[2 * r];
    hv_out[2 * r + 1] = hvl[2 * r + 1];
    if (tpl) {
      int best = 0;
      float bv = tpl[r * nt];
      for (int j = 1; j < nt; ++j) {
        float v = tpl[r * nt + j];
        if (v > bv) bv = v, best = j;
      }
      tp_out[r] = (float)best;
    }
  }
}

// UNetModel.infer_batch tail (unet.py:458-465): softmax over C logits -> bilinear x2 (align_corners=False) -> centre crop
// to (H, W) of the (2H, 2W) map -> NHWC probabilities.  logits fp32 [N, H, W, C]; out fp32 [N, H, W, C].
__global__ void __launch_bounds__(256) softmax_up2_crop_kernel(const float* __restrict__ lg, int Nimg, int H, int W, int C,
                                                               float* __restrict__ out) {
  const long long total = (long long)Nimg * H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int xo = (int)(i % W);
    const int yo = (int)((i / W) % H);
    const int img = (int)(i / ((long long)W * H));
    // output pixel (yo, xo) of the crop = pixel (yo + H/2, xo + W/2) of the 2x map
    const int Y = yo + H / 2, X = xo + W / 2;
    // align_corners=False: src = (dst + 0.5)/2 - 0.5, clamped at 0
    float sy = fmaxf((Y + 0.5f) * 0.5f - 0.5f, 0.f), sx = fmaxf((X + 0.5f) * 0.5f - 0.5f, 0.f);
    int y0 = (int)sy, x0 = (int)sx;
    int y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);
    const float fy = sy - (float)y0, fx = sx - (float)x0;
    float acc[8];
    for (int c = 0; c < C; ++c) acc[c] = 0.f;
    const int ys[2] = {y0, y1}, xs[2] = {x0, x1};
    const float wy[2] = {1.f - fy, fy}, wx[2] = {1.f - fx, fx};
    for (int a = 0; a < 2; ++a)
      for (int b = 0; b < 2; ++b) {
        const float* p = lg + ((long long)(img * H + ys[a]) * W + xs[b]) * C;
        float m = p[0];
        for (int c = 1; c < C; ++c) m = fmaxf(m, p[c]);
        float z = 0.f, e[8];
        for (int c = 0; c < C; ++c) {
          e[c] = expf(p[c] - m);
          z += e[c];
        }
        const float wgt = wy[a] * wx[b] / z;
        for (int c = 0; c < C; ++c) acc[c] = fmaf(wgt, e[c], acc[c]);
      }
    for (int c = 0; c < C; ++c) out[i * C + c] = acc[c];
  }
}

}  // namespace tiab

// =================================================================================================
// host side
// =================================================================================================
using namespace tiab;

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn(tiab_ctx* h) {
  static PFN_encodeTiled fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !p) {
    h->err = "cuTensorMapEncodeTiled driver entry point not available";
    return nullptr;
  }
  fn = reinterpret_cast<PFN_encodeTiled>(p);
  return fn;
}

// 2-D matrix [rows, cols] with row pitch ld (elements); box = {64 cols, box_rows}.
// raw == 0: fp16 operand tile, 128-byte swizzle (what the UMMA descriptors expect).
// raw == 1 / 2: un-swizzled fp16 / fp32 tile for the operand-transform warps (dense [box_rows][64] in shared memory).
static int make_map_2d(tiab_ctx* h, CUtensorMap* map, const void* base, long long rows, int cols, long long ld,
                       int box_rows, int raw = 0) {
  PFN_encodeTiled enc = get_encode_fn(h);
  if (!enc) return TIAB_E_CUDA;
  const int esz = raw == 2 ? 4 : 2;
  if (((uintptr_t)base & 15) || ((ld * esz) & 15)) {
    h->err = "conv: operand base/pitch must be 16-byte aligned";
    return TIAB_E_INVALID;
  }
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * esz};
  cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, raw == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim,
                   gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, raw ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    h->err = "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")";
    return TIAB_E_CUDA;
  }
  return TIAB_OK;
}

template <int BN, int NSEG, int AMODE>
static int launch_conv(tiab_ctx* h, const CUtensorMap& ta, const CUtensorMap& ta2, const CUtensorMap& tb, ConvParams& p,
                       cudaStream_t s) {
  using S = ConvSmem<BN, NSEG, AMODE>;
  static bool attr_set = false;
  if (!attr_set) {
    TIAB_CUDA_OK(h, cudaFuncSetAttribute(conv_gemm_kernel<BN, NSEG, AMODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
    attr_set = true;
  }
  p.m_tiles_pg = (p.group_rows + BLOCK_M - 1) / BLOCK_M;
  p.num_m_tiles = p.groups * p.m_tiles_pg;
  p.num_n_tiles = p.N / BN;
  int tiles = p.num_m_tiles * p.num_n_tiles;
  int grid = std::min(tiles, h->sm_count);
  conv_gemm_kernel<BN, NSEG, AMODE><<<grid, AMODE ? CONV_THREADS_T : CONV_THREADS, S::TOTAL, s>>>(ta, ta2, tb, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    h->err = std::string("conv_gemm launch: ") + cudaGetErrorString(e);
    return TIAB_E_CUDA;
  }
  h->last_launches++;
  return TIAB_OK;
}

#define AUX_LAUNCH_CHECK(h, what)                                                \
  do {                                                                           \
    cudaError_t _e = cudaGetLastError();                                         \
    if (_e != cudaSuccess) {                                                     \
      (h)->err = std::string(what " launch: ") + cudaGetErrorString(_e);         \
      return TIAB_E_CUDA;                                                        \
    }                                                                            \
    (h)->last_launches++;                                                        \
  } while (0)

static int aux_grid(tiab_ctx* h, long long total, int threads) {
  long long g = (total + threads - 1) / threads;
  long long cap = (long long)h->sm_count * 16;
  return (int)std::max<long long>(1, std::min(g, cap));
}

extern "C" {

int tiab_conv_gemm(tiab_handle_t h, const tiab_conv_desc_t* d, void* stream) {
  if (!h) return TIAB_E_INVALID;
  if (!d || !d->x || !d->w || d->M <= 0 || d->N <= 0 || d->taps <= 0 || d->taps > TIAB_MAX_TAPS || d->cin <= 0) {
    h->err = "tiab_conv_gemm: invalid descriptor";
    return TIAB_E_INVALID;
  }
  if (d->N % 32) {
    h->err = "tiab_conv_gemm: N (output channels) must be a multiple of 32";
    return TIAB_E_INVALID;
  }
  if (d->x_rows < 1 || d->Hin <= 0 || d->Win <= 0) {
    h->err = "tiab_conv_gemm: invalid grid";
    return TIAB_E_INVALID;
  }
  const int amode = (d->pre_scale || d->x_is_f32) ? 1 : 0;
  const int nseg = (amode ? d->split : (d->x_lo_off != 0)) ? 3 : 1;
  const int kchunks = (d->cin + BLOCK_K - 1) / BLOCK_K;
  const long long kcols = (long long)d->taps * (nseg == 3 ? 2 : 1) * kchunks * BLOCK_K;  // per tap: [W_hi | W_lo] in split mode
  if (d->w_ld < kcols) {
    h->err = "tiab_conv_gemm: weight pitch smaller than taps*planes*ceil64(cin)";
    return TIAB_E_INVALID;
  }
  TIAB_ON_DEVICE(h);
  const int BN = (d->N % 256 == 0) ? 256 : (d->N % 128 == 0) ? 128 : (d->N % 64 == 0) ? 64 : 32;
  CUtensorMap ta, ta2, tb;
  int xcols = d->cin;   // extent of the operand map along the channels (taps may address different channel blocks of a row)
  for (int t = 0; t < d->taps; ++t) {
    if (d->tap_coff[t] < 0 || (d->tap_coff[t] & 7) || (d->tap_coff[t] && (d->cin % BLOCK_K))) {
      h->err = "tiab_conv_gemm: tap_coff must be >= 0, a multiple of 8, and needs cin % 64 == 0";
      return TIAB_E_INVALID;
    }
    xcols = std::max(xcols, d->tap_coff[t] + d->cin);
  }
  if (d->s2d_c && (!d->out16 || d->out32 || d->out16b || (d->s2d_c & 3))) {
    h->err = "tiab_conv_gemm: space-to-depth store is for out16 only";
    return TIAB_E_INVALID;
  }
  const int groups = d->groups > 1 ? d->groups : 1;
  if (groups > 1 && (d->group_rows <= 0 || (long long)d->group_rows * groups != d->M || (d->cin & 3))) {
    h->err = "tiab_conv_gemm: grouped launch needs M == groups * group_rows and cin % 4 == 0";
    return TIAB_E_INVALID;
  }
  int rc = make_map_2d(h, &tb, d->w, (long long)groups * d->N, (int)kcols, d->w_ld, BN);
  if (rc) return rc;
  if (amode) {
    if ((d->cin & 3) || (d->x_lo_off & 7)) {
      h->err = "tiab_conv_gemm: transform mode needs cin % 4 == 0 and a 16-byte aligned lo plane";
      return TIAB_E_INVALID;
    }
    // raw tiles for the transform warps; the map's inner extent is the logical cin, so the K padding reads as zeros
    rc = make_map_2d(h, &ta, d->x, d->x_rows, xcols, d->x_ld, BLOCK_M, d->x_is_f32 ? 2 : 1);
    if (rc) return rc;
    ta2 = ta;
    if (!d->x_is_f32 && d->x_lo_off) {
      rc = make_map_2d(h, &ta2, (const __half*)d->x + d->x_lo_off, d->x_rows, xcols, d->x_ld, BLOCK_M, 1);
      if (rc) return rc;
    }
  } else {
    rc = make_map_2d(h, &ta, d->x, d->x_rows, xcols, d->x_ld, BLOCK_M);
    if (rc) return rc;
    rc = make_map_2d(h, &ta2, (const __half*)d->x + d->x_lo_off, d->x_rows, xcols, d->x_ld, BLOCK_M);
    if (rc) return rc;
  }
  ConvParams p;
  memset(&p, 0, sizeof(p));
  p.M = d->M;
  p.groups = groups, p.group_rows = groups > 1 ? d->group_rows : d->M, p.a_grouped = groups > 1 ? (d->x_grouped != 0) : 1;
  p.N = d->N;
  p.taps = d->taps;
  p.kchunks = kchunks;
  p.nseg = nseg;
  p.a_src = d->x, p.a_rows = d->x_rows, p.a_ld = (int)d->x_ld, p.a_lo = d->x_lo_off, p.a_is_f32 = d->x_is_f32, p.a_cin = d->cin;
  p.pre_scale = (const float*)d->pre_scale, p.pre_shift = (const float*)d->pre_shift, p.pre_relu = d->pre_relu;
  for (int t = 0; t < d->taps; ++t) p.tap_off[t] = d->tap_off[t], p.tap_coff[t] = d->tap_coff[t];
  p.s2d_c = d->s2d_c;
  p.oscale = (const float*)d->out_scale;
  if (d->head_w) {
    if (d->N != 64 || amode || groups > 4 || d->out16 || d->out32 || d->out16b || d->res32 || d->res16 || !d->head_b || !d->head_scale ||
        !d->head_shift) {
      h->err = "tiab_conv_gemm: the fused head needs N == 64, a TMA operand, <= 4 groups and no other output";
      return TIAB_E_INVALID;
    }
    p.head_w = (const float*)d->head_w, p.head_b = (const float*)d->head_b;
    p.head_scale = (const float*)d->head_scale, p.head_shift = (const float*)d->head_shift;
    for (int g = 0; g < groups; ++g) {
      if (!d->head_out[g] || d->head_nout[g] < 1 || d->head_nout[g] > 8 || d->head_ld[g] < d->head_nout[g]) {
        h->err = "tiab_conv_gemm: fused head: 1 <= nout <= 8 and an output per group";
        return TIAB_E_INVALID;
      }
      p.head_out[g] = (float*)d->head_out[g], p.head_nout[g] = d->head_nout[g], p.head_ld[g] = d->head_ld[g];
    }
  }
  p.Hin = d->Hin, p.Win = d->Win;
  p.skip = (d->y0 > 0 || d->y1 < d->Hin) ? 1 : 0;
  {
    const bool r32 = d->res32 != nullptr, r16 = d->res16 != nullptr, o32 = d->out32 != nullptr, o16 = d->out16 != nullptr,
               o16b = d->out16b != nullptr;
    p.epi_mode = 0;
    if (!r32 && !r16 && !o32 && o16 && !o16b) p.epi_mode = d->lo16 ? 1 : 2;
    else if (!r16 && o32 && !o16 && !o16b) p.epi_mode = r32 ? 4 : 3;
    else if (r32 && !o32 && !o16 && o16b && d->lo16b) p.epi_mode = 5;
  }
  p.y0 = d->y0, p.y1 = d->y1, p.x0 = d->x0, p.x1 = d->x1;
  p.Hd = d->Hd, p.Wd = d->Wd, p.dy = d->dy, p.dx = d->dx;
  p.Hr = d->Hr, p.Wr = d->Wr, p.ry = d->ry, p.rx = d->rx;
  p.out16 = (__half*)d->out16, p.ld16 = d->ld16, p.coff16 = d->coff16, p.lo16 = d->lo16;
  p.out32 = (float*)d->out32, p.ld32 = d->ld32, p.coff32 = d->coff32;
  p.out16b = (__half*)d->out16b, p.ld16b = d->ld16b, p.coff16b = d->coff16b, p.lo16b = d->lo16b;
  p.scale2 = (const float*)d->scale2, p.shift2 = (const float*)d->shift2;
  p.bias = (const float*)d->bias;
  p.res32 = (const float*)d->res32, p.res16 = (const __half*)d->res16, p.ldres = d->ldres, p.coffres = d->coffres;
  p.reslo = d->reslo;
  p.relu = d->relu;
  if (p.out16b && (!p.scale2 || !p.shift2)) {
    h->err = "tiab_conv_gemm: out16b needs scale2/shift2";
    return TIAB_E_INVALID;
  }
  cudaStream_t s = (cudaStream_t)stream;
#define TIAB_CONV_DISPATCH(NS, AM)                                          \
  switch (BN) {                                                             \
    case 256: return launch_conv<256, NS, AM>(h, ta, ta2, tb, p, s);        \
    case 128: return launch_conv<128, NS, AM>(h, ta, ta2, tb, p, s);        \
    case 64: return launch_conv<64, NS, AM>(h, ta, ta2, tb, p, s);          \
    default: return launch_conv<32, NS, AM>(h, ta, ta2, tb, p, s);          \
  }
  if (amode) {
    if (nseg == 3) { TIAB_CONV_DISPATCH(3, 1) }
    TIAB_CONV_DISPATCH(1, 1)
  }
  if (nseg == 3) { TIAB_CONV_DISPATCH(3, 0) }
  TIAB_CONV_DISPATCH(1, 0)
#undef TIAB_CONV_DISPATCH
}

int tiab_im2col(tiab_handle_t h, const void* src, int src_is_u8, float u8_scale, int src_lo, int N, int H, int W, int C,
                int ldsrc, int KH, int KW, int stride, int pad_t, int pad_l, int Ho, int Wo, void* out_f16, int ldout,
                int Kpad, int out_lo, void* stream) {
  if (!h || !src || !out_f16 || Kpad < KH * KW * C || ldout < Kpad) return TIAB_E_INVALID;
  TIAB_ON_DEVICE(h);
  if (!src_is_u8 && (C % 8) == 0 && (Kpad % 8) == 0 && (ldsrc % 8) == 0 && (ldout % 8) == 0 && (src_lo % 8) == 0 && (out_lo % 8) == 0 &&
      ((uintptr_t)src & 15) == 0 && ((uintptr_t)out_f16 & 15) == 0 && (!out_lo || src_lo)) {
    long long total8 = (long long)N * Ho * Wo * (Kpad >> 3);
    im2col_vec8_kernel<<<aux_grid(h, total8, 256), 256, 0, (cudaStream_t)stream>>>((const __half*)src, src_lo, N, H, W, C, ldsrc, KH, KW,
                                                                                  stride, pad_t, pad_l, Ho, Wo, (__half*)out_f16, ldout,
                                                                                  Kpad, out_lo);
    AUX_LAUNCH_CHECK(h, "im2col_vec8");
    return TIAB_OK;
  }
  if (src_is_u8 && C == 3 && ldsrc == 3 && (Kpad % 8) == 0 && (ldout % 8) == 0 && (out_lo % 8) == 0 && ((uintptr_t)out_f16 & 15) == 0 &&
      KH <= 15 && KW <= 15 && (long long)(KH * (long long)W + KW) * 3 < (1 << 24)) {
    long long t3 = (long long)N * Ho * Wo * (Kpad >> 3);
    im2col_u8c3_kernel<<<aux_grid(h, t3, 256), 256, Kpad * sizeof(uint32_t), (cudaStream_t)stream>>>((const uint8_t*)src, u8_scale, N, H, W, KH, KW, stride, pad_t,
                                                                             pad_l, Ho, Wo, (__half*)out_f16, ldout, Kpad, out_lo);
    AUX_LAUNCH_CHECK(h, "im2col_u8c3");
    return TIAB_OK;
  }
  long long total = (long long)N * Ho * Wo * Kpad;
  im2col_kernel<<<aux_grid(h, total, 256), 256, 0, (cudaStream_t)stream>>>(src, src_is_u8, u8_scale, src_lo, N, H, W, C, ldsrc,
                                                                           KH, KW, stride, pad_t, pad_l, Ho, Wo,
                                                                           (__half*)out_f16, ldout, Kpad, out_lo);
  AUX_LAUNCH_CHECK(h, "im2col");
  return TIAB_OK;
}

int tiab_bn_relu(tiab_handle_t h, const void* src, int src_is_f32, int ldsrc, int coffsrc, int src_lo, const float* scale,
                 const float* shift, int relu, long long rows, int C, void* dst_f16, int lddst, int coffdst, int dst_lo,
                 void* stream) {
  if (!h || !src || !dst_f16 || (C & 7) || (ldsrc & 7) || (lddst & 7) || (coffsrc & 7) || (coffdst & 7) || (src_lo & 7) || (dst_lo & 7))
    return TIAB_E_INVALID;
  TIAB_ON_DEVICE(h);
  bn_relu_kernel<<<aux_grid(h, rows * (C >> 3), 256), 256, 0, (cudaStream_t)stream>>>(src, src_is_f32, ldsrc, coffsrc, src_lo, scale,
                                                                                       shift, relu, rows, C, (__half*)dst_f16, lddst,
                                                                                       coffdst, dst_lo);
  AUX_LAUNCH_CHECK(h, "bn_relu");
  return TIAB_OK;
}

int tiab_upsample_add(tiab_handle_t h, const void* a_f16, int Ha, int Wa, int Hag, int Wag, int ay, int ax, int lda, int a_lo,
                      const void* b_f16, int Hb, int Wb, int by, int bx, int ldb, int b_lo, int N, int C, const float* scale,
                      const float* shift, int relu, void* dst_f16, int Hd, int Wd, int dy, int dx, int ldd, int dst_lo,
                      void* stream) {
  if (!h || !a_f16 || !dst_f16 || (C & 7) || (lda & 7) || (ldb & 7) || (ldd & 7)) return TIAB_E_INVALID;
  TIAB_ON_DEVICE(h);
  long long total = (long long)N * 4 * Ha * Wa * (C >> 3);
  upsample_add_kernel<<<aux_grid(h, total, 256), 256, 0, (cudaStream_t)stream>>>(
      (const __half*)a_f16, Ha, Wa, Hag, Wag, ay, ax, lda, a_lo, (const __half*)b_f16, Hb, Wb, by, bx, ldb, b_lo, N, C, scale, shift,
      relu, (__half*)dst_f16, Hd, Wd, dy, dx, ldd, dst_lo);
  AUX_LAUNCH_CHECK(h, "upsample_add");
  return TIAB_OK;
}

int tiab_maxpool3x3s2(tiab_handle_t h, const void* src_f16, int N, int H, int W, int C, int lds, int src_lo, void* dst_f16,
                      int Hd, int Wd, int dy, int dx, int ldd, int dst_lo, void* stream) {
  if (!h || !src_f16 || !dst_f16 || (C & 7) || (lds & 7) || (ldd & 7)) return TIAB_E_INVALID;
  TIAB_ON_DEVICE(h);
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  long long total = (long long)N * Ho * Wo * (C >> 3);
  maxpool3x3s2_kernel<<<aux_grid(h, total, 256), 256, 0, (cudaStream_t)stream>>>((const __half*)src_f16, N, H, W, C, lds, src_lo,
                                                                                 (__half*)dst_f16, Ho, Wo, Hd, Wd, dy, dx, ldd, dst_lo);
  AUX_LAUNCH_CHECK(h, "maxpool");
  return TIAB_OK;
}

int tiab_pointwise_head(tiab_handle_t h, const void* x_f16, long long rows, int C, int ldx, int x_lo, const float* scale,
                        const float* shift, int relu, const float* Wt, const float* bias, int nout, float* out, int ldout,
                        int coffout, void* stream) {
  if (!h || !x_f16 || !Wt || !out || nout < 1 || nout > 16 || (C & 7) || (ldx & 7) || (x_lo & 7)) return TIAB_E_INVALID;
  TIAB_ON_DEVICE(h);
  size_t sm = (size_t)(nout * C + 2 * C) * sizeof(float);
  pointwise_head_kernel<<<aux_grid(h, rows, 128), 128, sm, (cudaStream_t)stream>>>((const __half*)x_f16, rows, C, ldx, x_lo, scale, shift,
                                                                                    relu, Wt, bias, nout, out, ldout, coffout);
  AUX_LAUNCH_CHECK(h, "pointwise_head");
  return TIAB_OK;
}

int tiab_gap_fc_softmax(tiab_handle_t h, const void* x_f16, int N, int HW, int C, int ldx, int x_lo, const float* Wt,
                        const float* bias, int ncls, float* feat_out, float* probs, void* stream) {
  if (!h || !x_f16 || !Wt || !bias || !probs || ncls < 1 || ncls > 1024) return TIAB_E_INVALID;
  TIAB_ON_DEVICE(h);
  size_t sm = (size_t)(C + ncls) * sizeof(float);
  gap_fc_softmax_kernel<<<N, 256, sm, (cudaStream_t)stream>>>((const __half*)x_f16, HW, C, ldx, x_lo, Wt, bias, ncls, feat_out, probs);
  AUX_LAUNCH_CHECK(h, "gap_fc_softmax");
  return TIAB_OK;
}

int tiab_hovernet_heads(tiab_handle_t h, const float* np_logits, const float* hv_logits, const float* tp_logits, int nt,
                        long long rows, float* np_out, float* hv_out, float* tp_out, void* stream) {
  if (!h || !np_logits || !hv_logits || !np_out || !hv_out || (tp_logits && (!tp_out || nt < 1))) return TIAB_E_INVALID;
  TIAB_ON_DEVICE(h);
  hovernet_heads_kernel<<<aux_grid(h, rows, 256), 256, 0, (cudaStream_t)stream>>>(np_logits, hv_logits, tp_logits, nt, rows, np_out,
                                                                                  hv_out, tp_out);
  AUX_LAUNCH_CHECK(h, "hovernet_heads");
  return TIAB_OK;
}

int tiab_softmax_up2_crop(tiab_handle_t h, const float* logits, int N, int H, int W, int C, float* out, void* stream) {
  if (!h || !logits || !out || C < 1 || C > 8) return TIAB_E_INVALID;
  TIAB_ON_DEVICE(h);
  softmax_up2_crop_kernel<<<aux_grid(h, (long long)N * H * W, 256), 256, 0, (cudaStream_t)stream>>>(logits, N, H, W, C, out);
  AUX_LAUNCH_CHECK(h, "softmax_up2_crop");
  return TIAB_OK;
}

}  // extern "C"
