// contour_math.cuh — the outer border of ONE labelled instance, as cv2.findContours(mask, RETR_TREE, CHAIN_APPROX_SIMPLE)[0][0] returns
// it for a single 8-connected component (tiatoolbox/models/architecture/hovernet.py:685-693: get_instance_info traces every nucleus).
//
// OpenCV follows borders with Suzuki & Abe's algorithm (imgproc contours: the outer border starts at the component's first pixel in
// raster order, the first neighbour is searched clockwise starting from the west, every further one counter-clockwise from the
// direction the walk came from; CHAIN_APPROX_SIMPLE keeps a point only when the step direction changes).  Restated here from that
// published algorithm as a __host__ __device__ function so that the code that runs per instance on the GPU is fuzz-tested against cv2
// on the CPU (tests/test_oracle_contour.py) through tiab_dbg_trace_contour.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define TIAB_CT_HD __host__ __device__ __forceinline__
#else
#define TIAB_CT_HD inline
#endif

namespace tiab {

struct LabelView {   // an int32 label image; pixels outside it are background
  const int32_t* map;
  int H, W;
  int32_t id;
  TIAB_CT_HD bool fg(int x, int y) const { return x >= 0 && y >= 0 && x < W && y < H && map[(long long)y * W + x] == id; }
};

// Freeman codes as OpenCV numbers them: 0 = east, then counter-clockwise on the screen (1 = north-east, 2 = north, ... 7 = south-east)
TIAB_CT_HD int ct_dx(int s) { return (s == 0 || s == 1 || s == 7) ? 1 : ((s == 3 || s == 4 || s == 5) ? -1 : 0); }
TIAB_CT_HD int ct_dy(int s) { return (s == 1 || s == 2 || s == 3) ? -1 : ((s == 5 || s == 6 || s == 7) ? 1 : 0); }

// Trace the outer border of the component of `v.id` that contains (x0, y0), which must be its first pixel in raster order.
// Writes at most `cap` points (x, y interleaved) to `out` and returns the number of points of the CHAIN_APPROX_SIMPLE contour
// (also when it exceeds cap, so that a caller can size a second pass).
TIAB_CT_HD int trace_outer_contour(const LabelView& v, int x0, int y0, int* out, int cap) {
  int n = 0;
  // first neighbour: clockwise from the west (the west pixel itself is background by construction)
  int s = 4;
  const int s_end0 = 4;
  int x1 = x0, y1 = y0;
  do {
    s = (s - 1) & 7;
    x1 = x0 + ct_dx(s);
    y1 = y0 + ct_dy(s);
  } while (!v.fg(x1, y1) && s != s_end0);
  if (s == s_end0) {   // isolated pixel
    if (n < cap) out[2 * n] = x0, out[2 * n + 1] = y0;
    return 1;
  }
  int x3 = x0, y3 = y0;          // current border pixel
  int prev_s = s ^ 4;
  for (;;) {
    // next border pixel: counter-clockwise from the direction we came from
    int x4, y4;
    for (;;) {
      s = (s + 1) & 7;
      x4 = x3 + ct_dx(s);
      y4 = y3 + ct_dy(s);
      if (v.fg(x4, y4)) break;
    }
    if (s != prev_s) {           // CHAIN_APPROX_SIMPLE: a vertex where the direction changes
      if (n < cap) out[2 * n] = x3, out[2 * n + 1] = y3;
      ++n;
      prev_s = s;
    }
    if (x4 == x0 && y4 == y0 && x3 == x1 && y3 == y1) break;   // back at the start through the first edge
    x3 = x4;
    y3 = y4;
    s = (s + 4) & 7;
  }
  return n;
}

}  // namespace tiab
