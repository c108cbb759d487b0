// watershed_math.cuh — the sequential priority flood of skimage.segmentation.watershed (connectivity 1, no compactness, no watershed line),
// the last step of HoVerNet._proc_np_hv (tiatoolbox/models/architecture/hovernet.py:616), as __host__ __device__ code so that the flood the
// GPU runs per foreground component is checked against the oracle on the CPU (tiab_dbg_watershed, tests/test_oracle_watershed.py).
//
// Semantics (skimage/segmentation/_watershed_cy.pyx, restated in oracle/hovernet_postproc_oracle.py:watershed): markers enter the queue
// with age 0 in raveled order; the queue pops the smallest (value, age); the popped pixel's unlabelled in-mask 4-neighbours are visited
// in the order (-row, -1, +1, +row), labelled when PUSHED, each with the next age.  Ties of (value, age) can only occur between marker
// pixels; skimage leaves them to its heap's internals, the oracle (Python tuples) breaks them by the pixel index — the total order
// (value, age, index) used here makes the result independent of the queue's shape, so any d-ary heap gives the oracle's answer.
//
// Version 2 of the flood (r1h finding: on maps with ~40 nuclei the binary heap in global memory made post-processing as slow as the
// network): an 8-ary heap — one 128-byte line holds a node's children, the sift-down depth drops from ~12 to ~4 dependent steps.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define TIAB_WS_HD __host__ __device__ __forceinline__
#else
#define TIAB_WS_HD inline
#endif

namespace tiab {

struct WsItem {
  double v;
  unsigned age;
  unsigned idx;
};

TIAB_WS_HD bool ws_less(const WsItem& a, const WsItem& b) {
  if (a.v != b.v) return a.v < b.v;
  if (a.age != b.age) return a.age < b.age;
  return a.idx < b.idx;
}

template <int D>
TIAB_WS_HD void ws_push(WsItem* h, int& n, const WsItem e) {
  int c = n++;
  while (c > 0) {
    const int par = (c - 1) / D;
    if (!ws_less(e, h[par])) break;
    h[c] = h[par];
    c = par;
  }
  h[c] = e;
}

template <int D>
TIAB_WS_HD WsItem ws_pop(WsItem* h, int& n) {
  const WsItem top = h[0];
  const WsItem last = h[--n];
  int p = 0;
  for (;;) {
    const int c0 = D * p + 1;
    if (c0 >= n) break;
    const int c1 = c0 + D < n ? c0 + D : n;
    int best = c0;
    WsItem bv = h[c0];
    for (int c = c0 + 1; c < c1; ++c) {   // the children are contiguous: independent loads
      const WsItem cv = h[c];
      if (ws_less(cv, bv)) bv = cv, best = c;
    }
    if (!ws_less(bv, last)) break;
    h[p] = bv;
    p = best;
  }
  if (n > 0) h[p] = last;
  return top;
}

// Flood the pixels of one component.  comp == nullptr: every mask pixel of the [x0, x1] x [y0, y1] box may seed (whole-image flood);
// otherwise only pixels with comp[p] == r do (one connected foreground component: its in-mask neighbours belong to it by construction).
// `heap` must hold as many items as the component has pixels (every pixel is pushed at most once).  `out` must be 0 on the component.
template <int D>
TIAB_WS_HD void ws_flood(const int* comp, int r, int x0, int y0, int x1, int y1, const uint8_t* mask, const int* marker, const double* dist,
                         int H, int W, WsItem* heap, int* out) {
  int hn = 0;
  for (int y = y0; y <= y1; ++y)
    for (int x = x0; x <= x1; ++x) {
      const int p = y * W + x;
      if (comp ? comp[p] != r : !mask[p]) continue;
      const int m = marker[p];
      if (m > 0) {
        out[p] = m;
        ws_push<D>(heap, hn, WsItem{dist[p], 0u, (unsigned)p});
      }
    }
  unsigned age = 0;
  while (hn > 0) {
    const WsItem e = ws_pop<D>(heap, hn);
    const int p = (int)e.idx;
    const int x = p % W, y = p / W;
    const int l = out[p];
    const int nb[4] = {p - W, p - 1, p + 1, p + W};
    const bool ok[4] = {y > 0, x > 0, x < W - 1, y < H - 1};
    for (int k = 0; k < 4; ++k) {
      if (!ok[k]) continue;
      const int q = nb[k];
      if (!mask[q] || out[q]) continue;
      ++age;
      out[q] = l;
      ws_push<D>(heap, hn, WsItem{dist[q], age, (unsigned)q});
    }
  }
}

}  // namespace tiab
