// stitch.cu — WSI stitching on the device.
//
// Replaces the Python block loop of tiatoolbox/models/engine/semantic_segmentor.py:
//   merge_batch_to_canvas :1141-1183 (one patch row -> row canvas + uint8 count, all-zero blocks skipped and not counted),
//   merge_horizontal :1186-1263 (one canvas per patch row), merge_vertical_chunkwise :1398-1534 (seam rows of neighbouring row
//   canvases folded, count normalisation with 0 -> 1, clipping to the output shape).
//
// HBM-bound gather: one thread per output pixel sums the blocks that cover it IN THE REFERENCE'S ORDER (block index inside a
// patch row, then row by row), so float32 results are bit-identical to numpy's `canvas[...] += block`; no atomics.
// Algorithmic bytes per output pixel: 4*C*(1 + overlap factor) -- every block element is read once, every output written once.
// The x-axis of each patch row is cut at the block boundaries into segments with a fixed ordered block list (built on the host,
// O(#blocks log #blocks)); a pixel finds its segment by binary search.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/tiab200.h"
#include "ctx.h"

namespace tiab {

struct StitchRow {
  int y0, y1;        // global rows [y0, y1) covered by this patch row
  int bp_off, nseg;  // breakpoints bp[bp_off .. bp_off + nseg] ; segment s = [bp[s], bp[s+1])
};

// any(block != 0) per block (np.any: NaN counts, -0.0 does not)
__global__ void __launch_bounds__(256) stitch_block_flags(const float* __restrict__ blocks, long long per_block, int* __restrict__ flags) {
  const float* p = blocks + (long long)blockIdx.x * per_block;
  int any = 0;
  for (long long i = threadIdx.x; i < per_block && !any; i += blockDim.x) any |= (p[i] != 0.0f);
  any = __syncthreads_or(any);
  if (threadIdx.x == 0) flags[blockIdx.x] = any;
}

// sum (in block order) of the blocks of patch row `r` that cover column X at local row `ly`; returns the uint8-wrapped count
__device__ __forceinline__ unsigned row_sum(const float* __restrict__ blocks, const int* __restrict__ locs, const int* __restrict__ flags,
                                            const StitchRow& R, const int* __restrict__ bp, const int* __restrict__ seg_off,
                                            const int* __restrict__ seg_blocks, int X, int ly, int bh, int bw, int C, int row_width,
                                            float* acc) {
  for (int c = 0; c < C; ++c) acc[c] = 0.0f;
  if (X >= row_width || R.nseg == 0) return 0u;
  const int* b = bp + R.bp_off;
  if (X < b[0] || X >= b[R.nseg]) return 0u;
  int lo = 0, hi = R.nseg;   // largest s with b[s] <= X
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (b[mid] <= X) lo = mid; else hi = mid;
  }
  const int s = R.bp_off + lo;   // seg_off is indexed like bp (the last breakpoint of a row owns an empty list)
  unsigned cnt = 0;
  for (int k = seg_off[s]; k < seg_off[s + 1]; ++k) {
    const int i = seg_blocks[k];
    if (!flags[i]) continue;
    const int xs = locs[4 * i], ys = locs[4 * i + 1], ye = locs[4 * i + 3];
    if (ly >= ye - ys || ly >= bh) continue;            // canvas[0:ye-ys, ...] += block[0:ye-ys, ...]
    const float* p = blocks + (((long long)i * bh + ly) * bw + (X - xs)) * C;
    for (int c = 0; c < C; ++c) acc[c] = __fadd_rn(acc[c], p[c]);
    cnt = (cnt + 1u) & 0xFFu;
  }
  return cnt;
}

constexpr int STITCH_MAX_C = 16;

// mode 0: merge_batch_to_canvas (one row, canvas + count); mode 1: full stitch (rows folded + normalised)
__global__ void __launch_bounds__(256) stitch_gather(const float* __restrict__ blocks, const int* __restrict__ locs,
                                                     const int* __restrict__ flags, const StitchRow* __restrict__ rows, int nrows,
                                                     const int* __restrict__ bp, const int* __restrict__ seg_off,
                                                     const int* __restrict__ seg_blocks, int bh, int bw, int C, int row_width, int H,
                                                     int W, int mode, float* __restrict__ out, uint8_t* __restrict__ count) {
  const long long total = (long long)H * W;
  for (long long px = (long long)blockIdx.x * blockDim.x + threadIdx.x; px < total; px += (long long)gridDim.x * blockDim.x) {
    const int Y = (int)(px / W), X = (int)(px - (long long)Y * W);
    float acc[STITCH_MAX_C], tot[STITCH_MAX_C];
    if (mode == 0) {
      const unsigned cnt = row_sum(blocks, locs, flags, rows[0], bp, seg_off, seg_blocks, X, Y, bh, bw, C, row_width, acc);
      for (int c = 0; c < C; ++c) out[px * C + c] = acc[c];
      count[px] = (uint8_t)cnt;
      continue;
    }
    // rows are sorted by y0 and only neighbours overlap: r = last row with y0 <= Y, optionally preceded by r - 1
    const int Yg = Y + rows[0].y0;
    int lo = 0, hi = nrows;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (rows[mid].y0 <= Yg) lo = mid; else hi = mid;
    }
    unsigned cnt = 0;
    bool have = false;
    for (int c = 0; c < C; ++c) tot[c] = 0.0f;
    for (int r = max(lo - 1, 0); r <= lo; ++r) {
      const StitchRow R = rows[r];
      if (Yg < R.y0 || Yg >= R.y0 + bh) continue;   // a row canvas is bh rows tall (rows past ye-ys stay zero / uncounted)
      const unsigned c_r = row_sum(blocks, locs, flags, R, bp, seg_off, seg_blocks, X, Yg - R.y0, bh, bw, C, row_width, acc);
      if (!have) {
        for (int c = 0; c < C; ++c) tot[c] = acc[c];
        have = true;
      } else {
        for (int c = 0; c < C; ++c) tot[c] = __fadd_rn(tot[c], acc[c]);   // curr_chunk[-overlap:] += next_chunk[:overlap]
      }
      cnt = (cnt + c_r) & 0xFFu;
    }
    const float denom = (float)(cnt == 0 ? 1u : cnt);
    for (int c = 0; c < C; ++c) out[px * C + c] = __fdiv_rn(tot[c], denom);
  }
}

struct StitchTables {
  std::vector<StitchRow> rows;
  std::vector<int> bp, seg_off, seg_blocks;
};

// x-segments of the blocks [i0, i1): breakpoints + ordered covering lists (appended to t)
static void build_row(const int* locs, int i0, int i1, int row_width, int y0, int y1, StitchTables& t) {
  std::vector<int> b;
  for (int i = i0; i < i1; ++i) {
    const int xs = std::max(locs[4 * i], 0), xe = std::min(locs[4 * i + 2], row_width);
    if (xe > xs) b.push_back(xs), b.push_back(xe);
  }
  std::sort(b.begin(), b.end());
  b.erase(std::unique(b.begin(), b.end()), b.end());
  StitchRow R{y0, y1, (int)t.bp.size(), b.empty() ? 0 : (int)b.size() - 1};
  if (b.empty()) b.push_back(0);
  for (size_t s = 0; s < b.size(); ++s) {
    t.bp.push_back(b[s]);
    t.seg_off.push_back((int)t.seg_blocks.size());
    if (s + 1 < b.size())
      for (int i = i0; i < i1; ++i)
        if (locs[4 * i] <= b[s] && std::min(locs[4 * i + 2], row_width) >= b[s + 1]) t.seg_blocks.push_back(i);
  }
  t.rows.push_back(R);
}

static int run_stitch(tiab_ctx* h, const float* blocks, int N, int bh, int bw, int C, const int* locs, const StitchTables& t,
                      int row_width, int H, int W, int mode, float* out, uint8_t* count, cudaStream_t s) {
  const size_t nl = (size_t)std::max(N, 1) * 4, nr = t.rows.size(), nb = t.bp.size(), nsb = std::max<size_t>(t.seg_blocks.size(), 1);
  // one small staging allocation: locs | flags | rows | bp | seg_off (+1) | seg_blocks
  const size_t bytes = (nl + (size_t)std::max(N, 1) + nb + nb + 1 + nsb) * sizeof(int) + nr * sizeof(StitchRow) + 64;
  void* dev = nullptr;
  TIAB_CUDA_OK(h, cudaMallocAsync(&dev, bytes, s));
  int* d_locs = (int*)dev;
  int* d_flags = d_locs + nl;
  int* d_bp = d_flags + std::max(N, 1);
  int* d_off = d_bp + nb;
  int* d_sb = d_off + nb + 1;
  StitchRow* d_rows = (StitchRow*)(((uintptr_t)(d_sb + nsb) + 15) & ~(uintptr_t)15);
  std::vector<int> off = t.seg_off;
  off.push_back((int)t.seg_blocks.size());
  if (N) TIAB_CUDA_OK(h, cudaMemcpyAsync(d_locs, locs, (size_t)N * 4 * sizeof(int), cudaMemcpyHostToDevice, s));
  TIAB_CUDA_OK(h, cudaMemcpyAsync(d_bp, t.bp.data(), nb * sizeof(int), cudaMemcpyHostToDevice, s));
  TIAB_CUDA_OK(h, cudaMemcpyAsync(d_off, off.data(), off.size() * sizeof(int), cudaMemcpyHostToDevice, s));
  if (!t.seg_blocks.empty())
    TIAB_CUDA_OK(h, cudaMemcpyAsync(d_sb, t.seg_blocks.data(), t.seg_blocks.size() * sizeof(int), cudaMemcpyHostToDevice, s));
  TIAB_CUDA_OK(h, cudaMemcpyAsync(d_rows, t.rows.data(), nr * sizeof(StitchRow), cudaMemcpyHostToDevice, s));
  TIAB_CUDA_OK(h, cudaStreamSynchronize(s));   // the host vectors above go out of scope with the caller
  h->last_launches = 0;
  if (N) {
    stitch_block_flags<<<N, 256, 0, s>>>(blocks, (long long)bh * bw * C, d_flags);
    h->last_launches++;
  }
  const long long total = (long long)H * W;
  if (total > 0) {
    const int grid = (int)std::min<long long>((total + 255) / 256, (long long)h->sm_count * 16);
    stitch_gather<<<grid, 256, 0, s>>>(blocks, d_locs, d_flags, d_rows, (int)nr, d_bp, d_off, d_sb, bh, bw, C, row_width, H, W, mode, out,
                                       count);
    h->last_launches++;
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    h->err = std::string("stitch launch: ") + cudaGetErrorString(e);
    return TIAB_E_CUDA;
  }
  TIAB_CUDA_OK(h, cudaFreeAsync(dev, s));
  return TIAB_OK;
}

}  // namespace tiab

using namespace tiab;

extern "C" {

int tiab_merge_batch_to_canvas(tiab_handle_t h, const float* blocks_dev, int N, int bh, int bw, int C, const int* locs_host, int H, int W,
                               float* canvas_dev, uint8_t* count_dev, void* stream) {
  if (!h) return TIAB_E_INVALID;
  if (N < 0 || bh <= 0 || bw <= 0 || C <= 0 || C > STITCH_MAX_C || H < 0 || W < 0 || (N && (!blocks_dev || !locs_host)) ||
      ((long long)H * W > 0 && (!canvas_dev || !count_dev))) {
    h->err = "tiab_merge_batch_to_canvas: invalid argument (1 <= C <= 16)";
    return TIAB_E_INVALID;
  }
  for (int i = 0; i < N; ++i) {
    const int xs = locs_host[4 * i], ys = locs_host[4 * i + 1], xe = locs_host[4 * i + 2], ye = locs_host[4 * i + 3];
    if (xs < 0 || xe < xs || xe - xs > bw || ye < ys) {
      h->err = "tiab_merge_batch_to_canvas: a block needs 0 <= xs <= xe, xe - xs <= block width, ys <= ye";
      return TIAB_E_INVALID;
    }
  }
  TIAB_ON_DEVICE(h);
  StitchTables t;
  build_row(locs_host, 0, N, W, 0, H, t);
  return run_stitch(h, blocks_dev, N, bh, bw, C, locs_host, t, W, H, W, 0, canvas_dev, count_dev, (cudaStream_t)stream);
}

int tiab_stitch_patches(tiab_handle_t h, const float* blocks_dev, int N, int bh, int bw, int C, const int* locs_host, int canvas_width,
                        int out_H, int out_W, float* prob_dev, void* stream) {
  if (!h) return TIAB_E_INVALID;
  if (N <= 0 || bh <= 0 || bw <= 0 || C <= 0 || C > STITCH_MAX_C || !blocks_dev || !locs_host || out_H <= 0 || out_W <= 0 || !prob_dev ||
      canvas_width <= 0) {
    h->err = "tiab_stitch_patches: invalid argument (1 <= C <= 16, at least one block)";
    return TIAB_E_INVALID;
  }
  TIAB_ON_DEVICE(h);
  StitchTables t;
  int i0 = 0;
  for (int i = 1; i <= N; ++i) {
    if (i == N || locs_host[4 * i + 1] != locs_host[4 * i0 + 1]) {   // a new patch row starts when ys changes (infer_wsi :575-590)
      int y1 = locs_host[4 * i0 + 3];
      for (int k = i0; k < i; ++k) {
        const int xs = locs_host[4 * k], xe = locs_host[4 * k + 2];
        if (xs < 0 || xe < xs || std::min(xe, canvas_width) - xs > bw || locs_host[4 * k + 3] < locs_host[4 * k + 1]) {
          h->err = "tiab_stitch_patches: a block needs 0 <= xs <= xe, xe - xs <= block width, ys <= ye";
          return TIAB_E_INVALID;
        }
        y1 = std::max(y1, locs_host[4 * k + 3]);
      }
      build_row(locs_host, i0, i, canvas_width, locs_host[4 * i0 + 1], y1, t);
      i0 = i;
    }
  }
  // merge_vertical_chunkwise's assumptions: rows ascend, touch or overlap, and only neighbours overlap
  for (size_t r = 0; r + 1 < t.rows.size(); ++r) {
    if (t.rows[r + 1].y0 <= t.rows[r].y0 || t.rows[r].y1 < t.rows[r + 1].y0 || t.rows[r].y1 - t.rows[r].y0 != bh ||
        (r + 2 < t.rows.size() && t.rows[r].y0 + bh > t.rows[r + 2].y0)) {
      h->err = "tiab_stitch_patches: patch rows must be full height, ascend in y, touch or overlap, and overlap only their direct neighbours";
      return TIAB_E_INVALID;
    }
  }
  return run_stitch(h, blocks_dev, N, bh, bw, C, locs_host, t, canvas_width, out_H, out_W, 1, prob_dev, nullptr, (cudaStream_t)stream);
}

}  // extern "C"
