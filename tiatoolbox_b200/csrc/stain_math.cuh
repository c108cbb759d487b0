// stain_math.cuh — scalar ("finalise") steps of the stain path, shared by the fused per-patch
// kernel, the multi-pass general path and the host-side debug entry points (tiab_dbg_*).
// Everything here is float64 and runs on ONE thread per image between the pixel passes.
//
// Reference semantics (paths relative to /root/reference/tiatoolbox):
//   utils/misc.py:405-444   contrast_enhancer          -> percentile_u8 / contrast_lut
//   utils/misc.py:261-290   get_luminosity_tissue_mask -> cv2 8-bit Lab L via the fixed-point Y test
//   tools/stainextract.py:177-227 MacenkoExtractor     -> eigh3 / macenko_from_angles
//   tools/stainextract.py:281-322 VahadaneExtractor    -> lasso2 (sklearn LARS path) + vah_* steps
//   tools/stainnorm.py:49-113     get_concentrations / fit / transform -> pinv23 / apply matrix
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define TIAB_HD __host__ __device__ __forceinline__
#else
#define TIAB_HD inline
#endif

namespace tiab {

// ---- IEEE helpers that forbid FMA contraction (bit-exactness of the integer-producing steps) ----
TIAB_HD double dmul(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dmul_rn(a, b);
#else
  volatile double r = a * b;
  return r;
#endif
}
TIAB_HD double dadd(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dadd_rn(a, b);
#else
  volatile double r = a + b;
  return r;
#endif
}
TIAB_HD double dsub(double a, double b) { return dadd(a, -b); }
TIAB_HD double ddiv(double a, double b) {
#ifdef __CUDA_ARCH__
  return __ddiv_rn(a, b);
#else
  volatile double r = a / b;
  return r;
#endif
}

// numpy.lib._function_base_impl._lerp
TIAB_HD double np_lerp(double a, double b, double t) {
  double diff = dsub(b, a);
  double r = dadd(a, dmul(diff, t));
  if (t >= 0.5) r = dsub(b, dmul(diff, dsub(1.0, t)));
  return r;
}

// np.percentile(values, 100*q) ("linear") from an exact 256-bin histogram of uint8 values.
// n = total count.  q is the FRACTION (numpy divides the percentile by 100 first).
TIAB_HD double percentile_u8(const unsigned* hist, unsigned n, double q) {
  double vidx = dmul((double)(n - 1), q);
  double prev = floor(vidx);
  double gamma = dsub(vidx, prev);
  unsigned long long k = (unsigned long long)prev;
  unsigned long long k1 = k + 1 < n ? k + 1 : n - 1;
  unsigned long long cum = 0;
  int a = -1, b = -1;
  for (int v = 0; v < 256; ++v) {
    cum += hist[v];
    if (a < 0 && cum > k) a = v;
    if (b < 0 && cum > k1) { b = v; break; }
  }
  if (a < 0) a = 255;
  if (b < 0) b = a;
  return np_lerp((double)a, (double)b, gamma);
}

// contrast_enhancer's per-value map (utils/misc.py:431-444), given p2/p98 or min/max fallback.
TIAB_HD void contrast_range(const unsigned* hist, unsigned n, double q_lo, double q_hi, double* lo, double* hi) {
  double p_lo = percentile_u8(hist, n, q_lo);
  double p_hi = percentile_u8(hist, n, q_hi);
  if (p_lo >= p_hi) {
    int mn = 0, mx = 255;
    while (mn < 255 && hist[mn] == 0) ++mn;
    while (mx > 0 && hist[mx] == 0) --mx;
    p_lo = (double)mn;
    p_hi = (double)mx;
  }
  *lo = p_lo;
  *hi = p_hi;
}
TIAB_HD unsigned contrast_value(int v, double p_lo, double p_hi) {
  if (!(p_hi > p_lo)) return (unsigned)v;
  double x = (double)v;
  x = x < p_lo ? p_lo : (x > p_hi ? p_hi : x);  // np.clip
  x = ddiv(dsub(x, p_lo), dsub(p_hi, p_lo));
  x = dadd(dmul(x, 255.0), 0.0);
  return (unsigned)x;  // astype(uint8): truncation of a value in [0,255]
}

// sRGB gamma table of cv2's 8-bit RGB2LAB fixed-point path: rint(2040*g(i/255)).
TIAB_HD unsigned srgb_gamma_2040(int i) {
  double x = (double)i / 255.0;
  double g = x <= 0.04045 ? x / 12.92 : pow((x + 0.055) / 1.055, 2.4);
  return (unsigned)rint(2040.0 * g);
}

// Optical density of a uint8 level (utils/transforms.py:229-231): zeros count as ones.
TIAB_HD double od_of_level(int v) {
  if (v < 1) v = 1;
  double od = -log((double)v / 255.0);
  return od > 1e-6 ? od : 1e-6;
}

// ---- 3x3 symmetric eigen-decomposition (cyclic Jacobi, float64).  evals ascending, evecs[:,k]
//      stored column-wise as evecs[r*3+k] (same layout as numpy.linalg.eigh's return).
TIAB_HD void eigh3(const double a_in[9], double evals[3], double evecs[9]) {
  double a[3][3], v[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      a[i][j] = a_in[i * 3 + j];
      v[i][j] = i == j ? 1.0 : 0.0;
    }
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = fabs(a[0][1]) + fabs(a[0][2]) + fabs(a[1][2]);
    double diag = fabs(a[0][0]) + fabs(a[1][1]) + fabs(a[2][2]);
    if (off <= 1e-300 || off <= 1e-18 * diag) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        if (a[p][q] == 0.0) continue;
        double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
        double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 3; ++k) {  // A <- A * J
          double akp = a[k][p], akq = a[k][q];
          a[k][p] = c * akp - s * akq;
          a[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; ++k) {  // A <- J^T * A
          double apk = a[p][k], aqk = a[q][k];
          a[p][k] = c * apk - s * aqk;
          a[q][k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 3; ++k) {
          double vkp = v[k][p], vkq = v[k][q];
          v[k][p] = c * vkp - s * vkq;
          v[k][q] = s * vkp + c * vkq;
        }
      }
  }
  int idx[3] = {0, 1, 2};
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 2 - i; ++j)
      if (a[idx[j]][idx[j]] > a[idx[j + 1]][idx[j + 1]]) {
        int t = idx[j];
        idx[j] = idx[j + 1];
        idx[j + 1] = t;
      }
  for (int k = 0; k < 3; ++k) {
    evals[k] = a[idx[k]][idx[k]];
    for (int r = 0; r < 3; ++r) evecs[r * 3 + k] = v[r][idx[k]];
  }
}

// Macenko: covariance (ddof=1) of the tissue OD rows from raw moments
// mom = {n, sx, sy, sz, sxx, sxy, sxz, syy, syz, szz}; then the two principal eigenvectors,
// pointed so that their first component is >= 0 (stainextract.py:202-208).  E[ch*2+k].
TIAB_HD void macenko_basis(const double mom[10], double E[6]) {
  double n = mom[0];
  double s[3] = {mom[1], mom[2], mom[3]};
  double ss[3][3] = {{mom[4], mom[5], mom[6]}, {mom[5], mom[7], mom[8]}, {mom[6], mom[8], mom[9]}};
  double cov[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) cov[i * 3 + j] = (ss[i][j] - s[i] * s[j] / n) / (n - 1.0);
  double ev[3], vec[9];
  eigh3(cov, ev, vec);
  for (int ch = 0; ch < 3; ++ch) {
    E[ch * 2 + 0] = vec[ch * 3 + 2];
    E[ch * 2 + 1] = vec[ch * 3 + 1];
  }
  for (int k = 0; k < 2; ++k)
    if (E[0 * 2 + k] < 0)
      for (int ch = 0; ch < 3; ++ch) E[ch * 2 + k] = -E[ch * 2 + k];
}

// Pseudo-angle: a strictly monotone surrogate of atan2(y, x) on (-pi, pi], range [-2, 2].
// The order statistics of phi are taken on this key (no atan2 per pixel), then mapped back.
TIAB_HD float pseudo_angle(float x, float y) {
  float s = fabsf(x) + fabsf(y);
  float r = s > 0.f ? y / s : 0.f;
  if (x < 0.f) r = (y >= 0.f ? 2.f : -2.f) - r;
  return r;
}
TIAB_HD double angle_of_pseudo(double r) {
  if (r > 1.0) {
    double rr = 2.0 - r;  // y >= 0, x < 0
    return atan2(rr, -(1.0 - fabs(rr)));
  }
  if (r < -1.0) {
    double rr = -2.0 - r;
    return atan2(rr, -(1.0 - fabs(rr)));
  }
  return atan2(r, 1.0 - fabs(r));
}

// stainextract.py:221-227: the two stain vectors from the extreme angles, H first, rows unit-norm.
TIAB_HD void macenko_from_angles(const double E[6], double phi_min, double phi_max, double M[6]) {
  double v1[3], v2[3];
  double c1 = cos(phi_min), s1 = sin(phi_min), c2 = cos(phi_max), s2 = sin(phi_max);
  for (int ch = 0; ch < 3; ++ch) {
    v1[ch] = E[ch * 2] * c1 + E[ch * 2 + 1] * s1;
    v2[ch] = E[ch * 2] * c2 + E[ch * 2 + 1] * s2;
  }
  const double* r0 = v1[0] > v2[0] ? v1 : v2;
  const double* r1 = v1[0] > v2[0] ? v2 : v1;
  double n0 = sqrt(r0[0] * r0[0] + r0[1] * r0[1] + r0[2] * r0[2]);
  double n1 = sqrt(r1[0] * r1[0] + r1[1] * r1[1] + r1[2] * r1[2]);
  for (int ch = 0; ch < 3; ++ch) {
    M[ch] = r0[ch] / n0;
    M[3 + ch] = r1[ch] / n1;
  }
}

// get_concentrations (stainnorm.py:64-66): lstsq(M.T, od.T) == od @ pinv(M), pinv(M) = M^T (M M^T)^-1.
// A[ch*2+j].  Returns false when M is rank deficient.
TIAB_HD bool pinv23(const double M[6], double A[6]) {
  double g00 = M[0] * M[0] + M[1] * M[1] + M[2] * M[2];
  double g01 = M[0] * M[3] + M[1] * M[4] + M[2] * M[5];
  double g11 = M[3] * M[3] + M[4] * M[4] + M[5] * M[5];
  double det = g00 * g11 - g01 * g01;
  if (!(fabs(det) > 1e-14 * (g00 * g11 + 1e-300))) return false;
  double i00 = g11 / det, i01 = -g01 / det, i11 = g00 / det;
  for (int ch = 0; ch < 3; ++ch) {
    A[ch * 2 + 0] = M[ch] * i00 + M[3 + ch] * i01;
    A[ch * 2 + 1] = M[ch] * i01 + M[3 + ch] * i11;
  }
  return true;
}

// ---- scikit-learn LARS-lasso path for 2 variables (sklearn/linear_model/_least_angle.py
//      _lars_path_solver, method="lasso", Gram precomputed, alpha_min = reg/n, n_samples = n),
//      including the float32-eps early-stopping tolerance and the lasso drop step.
TIAB_HD double min_pos2(double a, double b) {
  const double big = 1.7976931348623157e308;
  double r = big;
  if (a > 0 && a < r) r = a;
  if (b > 0 && b < r) r = b;
  return r;
}
TIAB_HD void lasso2(const double G[4], const double c[2], double reg, double n, double out[2]) {
  const double eps32 = 1.1920928955078125e-07, tiny32 = 1.1754943508222875e-38;
  const double big = 1.7976931348623157e308;
  double alpha_min = reg / n;
  double coef[2] = {0, 0}, prev_coef[2] = {0, 0}, prev_alpha = 0;
  bool active[2] = {false, false};
  double sign[2] = {0, 0};
  int n_active = 0, n_iter = 0;
  bool drop = false;
  for (int it = 0; it < 16; ++it) {
    double corr[2] = {c[0] - (G[0] * coef[0] + G[1] * coef[1]), c[1] - (G[2] * coef[0] + G[3] * coef[1])};
    int jstar = -1;
    double cmax = 0;
    for (int j = 0; j < 2; ++j)
      if (!active[j] && (jstar < 0 || fabs(corr[j]) > cmax)) {
        jstar = j;
        cmax = fabs(corr[j]);
      }
    double alpha = cmax / n;
    if (alpha <= alpha_min + eps32) {
      if (fabs(alpha - alpha_min) > eps32 && n_iter > 0) {
        double ss = (prev_alpha - alpha_min) / (prev_alpha - alpha);
        for (int j = 0; j < 2; ++j) coef[j] = prev_coef[j] + ss * (coef[j] - prev_coef[j]);
      }
      break;
    }
    if (n_active >= 2) break;
    if (!drop) {
      sign[jstar] = corr[jstar] > 0 ? 1.0 : (corr[jstar] < 0 ? -1.0 : 0.0);
      active[jstar] = true;
      ++n_active;
    }
    double ls[2] = {0, 0}, aa;
    if (n_active == 1) {
      int a = active[0] ? 0 : 1;
      ls[a] = sign[a] / G[a * 3];
      aa = 1.0 / sqrt(ls[a] * sign[a]);
    } else {
      double det = G[0] * G[3] - G[1] * G[2];
      ls[0] = (G[3] * sign[0] - G[1] * sign[1]) / det;
      ls[1] = (G[0] * sign[1] - G[2] * sign[0]) / det;
      aa = 1.0 / sqrt(ls[0] * sign[0] + ls[1] * sign[1]);
    }
    ls[0] *= aa;
    ls[1] *= aa;
    double g1 = big, g2 = big;
    for (int j = 0; j < 2; ++j)
      if (!active[j]) {
        double ced = G[j * 2] * ls[0] + G[j * 2 + 1] * ls[1];
        ced = rint(ced * 1e15) / 1e15;  // np.around(..., decimals=15)
        g1 = min_pos2(g1, (cmax - corr[j]) / (aa - ced + tiny32));
        g2 = min_pos2(g2, (cmax + corr[j]) / (aa + ced + tiny32));
      }
    double gamma = g1 < g2 ? g1 : g2;
    if (cmax / aa < gamma) gamma = cmax / aa;
    drop = false;
    double z[2] = {-1, -1};
    for (int j = 0; j < 2; ++j)
      if (active[j]) z[j] = -coef[j] / (ls[j] + tiny32);
    double zpos = min_pos2(z[0], z[1]);
    bool dropj[2] = {false, false};
    if (zpos < gamma) {
      for (int j = 0; j < 2; ++j)
        if (active[j] && z[j] == zpos) {
          dropj[j] = true;
          sign[j] = -sign[j];
        }
      gamma = zpos;
      drop = true;
    }
    ++n_iter;
    prev_coef[0] = coef[0];
    prev_coef[1] = coef[1];
    prev_alpha = alpha;
    for (int j = 0; j < 2; ++j) coef[j] = active[j] ? prev_coef[j] + gamma * ls[j] : 0.0;
    for (int j = 0; j < 2; ++j)
      if (dropj[j]) {
        active[j] = false;
        sign[j] = 0;
        --n_active;
      }
  }
  out[0] = coef[0];
  out[1] = coef[1];
}

// ---- Vahadane dictionary learning (sklearn _dict_learning, n_components=2, max_iter=3) driven
//      by N-length reductions only.  Atom values at a pixel x (its 3 OD channels) are *replayed*:
//        d_k = u_k . x                                   (SVD init, D = U[:, :2]^T X)
//        step s on atom k:  d_k = max(x . g_s - beta_s * d_{1-k}, 0) * inv_s
//      because  D[k] += (B[:,k] - A[k] @ D) / A[k,k]  cancels D[k]'s own previous value.
struct VahState {
  double U[6];        // U[ch*2+k]: initial code (3x2), svd_flip'ed
  double g[4][3];     // per step
  double beta[4];
  double inv[4];      // 1 / max(||d_k||, 1) (set after the norm pass of that step)
  int nsteps;         // steps fully defined (incl. inv)
  double code[6];     // current code (3x2) code[ch*2+k]
  double Acc[4];      // code^T code of the current iteration
  double cost_prev;
  double trM;         // ||X||_F^2
  int iter;           // completed dictionary updates (0..2)
  int done;           // 1 => code is final
};

// X X^T (uncentred 2nd moments of the tissue OD rows) -> U (descending eigenvalues, svd_flip) and
// the first sparse code, for which G = U^T M U and C = U^T M need no pixel pass.
TIAB_HD void vah_encode(VahState* st, const double G[4], const double Cm[6], double reg, double n) {
  for (int ch = 0; ch < 3; ++ch) {
    double c2[2] = {Cm[0 * 3 + ch], Cm[1 * 3 + ch]};
    double w[2];
    lasso2(G, c2, reg, n, w);
    st->code[ch * 2 + 0] = w[0];
    st->code[ch * 2 + 1] = w[1];
  }
  for (int k = 0; k < 2; ++k)
    for (int l = 0; l < 2; ++l) {
      double s = 0;
      for (int ch = 0; ch < 3; ++ch) s += st->code[ch * 2 + k] * st->code[ch * 2 + l];
      st->Acc[k * 2 + l] = s;
    }
}
// returns false if the atom is unused (A[k,k] <= 1e-6): sklearn would draw a random atom.
TIAB_HD bool vah_begin_step(VahState* st, int k) {
  int s = st->nsteps;
  double akk = st->Acc[k * 2 + k];
  if (!(akk > 1e-6)) return false;
  for (int ch = 0; ch < 3; ++ch) st->g[s][ch] = st->code[ch * 2 + k] / akk;
  st->beta[s] = st->Acc[k * 2 + (1 - k)] / akk;
  st->inv[s] = 1.0;
  return true;
}
TIAB_HD void vah_end_step(VahState* st, double sumsq) {
  double nrm = sqrt(sumsq);
  st->inv[st->nsteps] = 1.0 / (nrm > 1.0 ? nrm : 1.0);
  st->nsteps += 1;
}
TIAB_HD void vah_init(VahState* st, const double mom[10], double reg) {
  double M[9] = {mom[4], mom[5], mom[6], mom[5], mom[7], mom[8], mom[6], mom[8], mom[9]};
  double ev[3], vec[9];
  eigh3(M, ev, vec);
  // descending order: columns 2,1 ; svd_flip (u-based): largest |entry| of each column positive
  for (int k = 0; k < 2; ++k) {
    int col = 2 - k;
    int im = 0;
    for (int r = 1; r < 3; ++r)
      if (fabs(vec[r * 3 + col]) > fabs(vec[im * 3 + col])) im = r;
    double sg = vec[im * 3 + col] < 0 ? -1.0 : 1.0;
    for (int r = 0; r < 3; ++r) st->U[r * 2 + k] = sg * vec[r * 3 + col];
  }
  st->nsteps = 0;
  st->iter = 0;
  st->done = 0;
  st->cost_prev = 0;
  st->trM = mom[4] + mom[7] + mom[9];
  // G = U^T M U, Cm = U^T M
  double MU[6];
  for (int r = 0; r < 3; ++r)
    for (int k = 0; k < 2; ++k) {
      double s = 0;
      for (int c = 0; c < 3; ++c) s += M[r * 3 + c] * st->U[c * 2 + k];
      MU[r * 2 + k] = s;
    }
  double G[4], Cm[6];
  for (int k = 0; k < 2; ++k) {
    for (int l = 0; l < 2; ++l) {
      double s = 0;
      for (int r = 0; r < 3; ++r) s += st->U[r * 2 + k] * MU[r * 2 + l];
      G[k * 2 + l] = s;
    }
    for (int ch = 0; ch < 3; ++ch) Cm[k * 3 + ch] = MU[ch * 2 + k];
  }
  vah_encode(st, G, Cm, reg, mom[0]);
}
// After a (G, Cm) pass over the dictionary produced by the just-finished update: cost of that
// iteration, convergence test (tol 1e-8 from the 2nd iteration) and the next sparse code.
// gc = {G00, G01, G11, C0r, C0g, C0b, C1r, C1g, C1b}
TIAB_HD void vah_after_gc(VahState* st, const double gc[9], double reg, double n) {
  double G[4] = {gc[0], gc[1], gc[1], gc[2]};
  double Cm[6] = {gc[3], gc[4], gc[5], gc[6], gc[7], gc[8]};
  double cross = 0, quad = 0, l1 = 0;
  for (int ch = 0; ch < 3; ++ch)
    for (int k = 0; k < 2; ++k) {
      cross += st->code[ch * 2 + k] * Cm[k * 3 + ch];
      l1 += fabs(st->code[ch * 2 + k]);
    }
  for (int k = 0; k < 2; ++k)
    for (int l = 0; l < 2; ++l) quad += st->Acc[k * 2 + l] * G[k * 2 + l];
  double cost = 0.5 * (st->trM - 2.0 * cross + quad) + reg * l1;
  st->iter += 1;
  if (st->iter >= 2) {
    double dE = st->cost_prev - cost;
    if (dE < 1e-8 * cost) {
      st->done = 1;  // sklearn breaks: the code of this iteration is returned
      return;
    }
  }
  st->cost_prev = cost;
  vah_encode(st, G, Cm, reg, n);
  if (st->iter >= 2) st->done = 1;  // 3rd encode is the last thing fit_transform uses
}
// stainextract.py:316-322: dictionary = code.T ; H row = larger first element ; unit rows.
TIAB_HD void vah_matrix(const VahState* st, double M[6]) {
  double r0[3], r1[3];
  for (int ch = 0; ch < 3; ++ch) {
    r0[ch] = st->code[ch * 2 + 0];
    r1[ch] = st->code[ch * 2 + 1];
  }
  const double* a = r0[0] < r1[0] ? r1 : r0;
  const double* b = r0[0] < r1[0] ? r0 : r1;
  double na = sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
  double nb = sqrt(b[0] * b[0] + b[1] * b[1] + b[2] * b[2]);
  for (int ch = 0; ch < 3; ++ch) {
    M[ch] = a[ch] / na;
    M[3 + ch] = b[ch] / nb;
  }
}

}  // namespace tiab
