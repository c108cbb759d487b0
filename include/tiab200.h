/* tiab200.h — C ABI of libtiab200.so: the B200-native per-patch hot path behind tiatoolbox's
 * Python seams.
 *
 * The reference (tiatoolbox 2.0.1, pure Python) has no FFI of its own; the seams this ABI sits
 * behind are duck-typed Python methods.  Each entry point names the reference
 * interface it replaces (paths relative to /root/reference/tiatoolbox).  INTEGRATION.md shows the
 * ctypes stub a tiatoolbox maintainer would add.
 *
 * Conventions: plain C types only; every function returns 0 on success or a negative TIAB_E_*;
 * tiab_last_error() gives the message; pointers named *_dev are device pointers, *_host host
 * pointers; `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Calls taking
 * only device pointers are asynchronous on `stream`; calls with host outputs synchronise it.
 * A handle owns one device's workspaces and is not thread-safe: use one handle per thread.
 */
#ifndef TIAB200_H
#define TIAB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tiab_ctx* tiab_handle_t;

enum {
  TIAB_OK = 0,
  TIAB_E_INVALID = -1,   /* bad argument */
  TIAB_E_CUDA = -2,      /* CUDA runtime error (see tiab_last_error) */
  TIAB_E_NOMEM = -3,
  TIAB_E_UNSUPPORTED = -4
};

/* per-image status written by the stain calls */
enum {
  TIAB_STAIN_OK = 0,
  TIAB_STAIN_EMPTY_TISSUE_MASK = 1, /* reference raises ValueError("Empty tissue mask computed.") utils/misc.py:286-288 */
  TIAB_STAIN_DEGENERATE = 2,        /* <2 tissue pixels / singular stain matrix: reference yields NaN */
  TIAB_STAIN_UNUSED_ATOM = 3        /* Vahadane: sklearn would re-sample a random atom (unspecified) */
};

enum { TIAB_STAIN_MACENKO = 0, TIAB_STAIN_VAHADANE = 1, TIAB_STAIN_RUIFROK = 2, TIAB_STAIN_CUSTOM = 3 };

int tiab_create(int device, tiab_handle_t* out);
void tiab_destroy(tiab_handle_t h);
const char* tiab_last_error(tiab_handle_t h);
/* "major.minor" of the ABI */
const char* tiab_version(void);

/* ---------------------------------------------------------------------------------------------
 * Stain normalisation  (tools/stainnorm.py:19-219, tools/stainextract.py:71-322,
 * utils/misc.py:261-290,405-444, utils/transforms.py:209-256)
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  int method;                  /* TIAB_STAIN_* — which extractor (stainextract.py:71,104,141,230) */
  int y_threshold;             /* tissue <=> cv2-Lab fixed-point Y <= y_threshold (1146 for luminosity 0.8; misc.py:281-283) */
  double angular_percentile;   /* MacenkoExtractor(angular_percentile=99)  stainextract.py:168-175 */
  double regularizer;          /* VahadaneExtractor(regularizer=0.1)       stainextract.py:259-279 */
  double fixed_matrix[6];      /* RUIFROK / CUSTOM: the constant 2x3 source stain matrix (row-major) */
} tiab_stain_params_t;

/* StainNormalizer.fit  (stainnorm.py:68-87): stain matrix + 99th-percentile concentrations of the
 * target.  target_dev: H*W*3 uint8 (HWC) on the device.  Outputs on the host; synchronises. */
int tiab_stain_fit(tiab_handle_t h, const tiab_stain_params_t* p, const uint8_t* target_dev, int H, int W,
                   double M_target_host[6], double maxC_target_host[2], int* status_host, void* stream);

/* StainNormalizer.transform over a batch (stainnorm.py:89-113 applied per image).
 * in_dev/out_dev: N*H*W*3 uint8 (NHWC), may alias.  Optional device outputs (NULL to skip):
 *   M_src_dev   N*6 doubles  — extractor.get_stain_matrix per image (stainextract.py:177/281)
 *   maxC_src_dev N*2 doubles — np.percentile(C, 99, axis=0) per image (stainnorm.py:103)
 *   status_dev  N ints       — TIAB_STAIN_*; on a non-OK status the image is copied through unchanged
 *   out_f32_dev N*H*W*3 float — the pre-quantisation value 255*exp(-C.M) clipped to [0,255]
 * Asynchronous on `stream`. */
int tiab_stain_transform_batch(tiab_handle_t h, const tiab_stain_params_t* p, const uint8_t* in_dev,
                               uint8_t* out_dev, int N, int H, int W, const double M_target[6],
                               const double maxC_target[2], double* M_src_dev, double* maxC_src_dev,
                               int* status_dev, float* out_f32_dev, void* stream);

/* Same with HOST buffers — the call the Python StainNormalizer.transform makes for numpy input:
 * pinned staging, H2D, kernels, D2H, synchronise.  M_src_host (N*6), maxC_src_host (N*2) and
 * status_host (N) are optional. */
int tiab_stain_transform_host(tiab_handle_t h, const tiab_stain_params_t* p, const uint8_t* in_host,
                              uint8_t* out_host, int N, int H, int W, const double M_target[6],
                              const double maxC_target[2], double* M_src_host, double* maxC_src_host,
                              int* status_host);

/* fit with a HOST target buffer (the call StainNormalizer.fit makes for a numpy target). */
int tiab_stain_fit_host(tiab_handle_t h, const tiab_stain_params_t* p, const uint8_t* target_host, int H, int W,
                        double M_target_host[6], double maxC_target_host[2], int* status_host);

/* get_luminosity_tissue_mask (utils/misc.py:261-290) for a batch: mask_dev N*H*W bytes (0/1). */
int tiab_stain_tissue_mask(tiab_handle_t h, const tiab_stain_params_t* p, const uint8_t* in_dev, uint8_t* mask_dev,
                           int N, int H, int W, int* status_dev, void* stream);

/* StainNormalizer.get_concentrations (stainnorm.py:49-66) for N images with one stain matrix:
 * conc_dev N*H*W*2 float32 = OD @ pinv(M). */
int tiab_stain_concentrations(tiab_handle_t h, const uint8_t* in_dev, int N, int H, int W, const double M[6],
                              float* conc_dev, void* stream);

/* StainAugmentor.augment (tools/stainaugment.py:171-206): out = uint8(clip(255 * exp(-(C' . M)), 0, 255)) with
 * C'[:, i] = C[:, i] * alpha[i] + beta[i] on the pixels where mask_dev != 0 (mask_dev = NULL: every pixel, `augment_background=True`).
 * conc_dev npix*2 float32 (tiab_stain_concentrations), mask_dev npix bytes (tiab_stain_tissue_mask), out_dev npix*3 uint8. */
int tiab_stain_augment(tiab_handle_t h, const float* conc_dev, const uint8_t* mask_dev, long long npix, const double M[6],
                       const double alpha[2], const double beta[2], uint8_t* out_dev, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Reinhard colour normalisation  (tools/stainnorm.py:222-367) and OpenCV's 8-bit RGB <-> Lab conversion it is built on
 * (cv2.cvtColor(COLOR_RGB2LAB / COLOR_LAB2RGB) on uint8: stainnorm.py:305, 334), as integer fixed-point device code that
 * equals cv2 on every one of the 2^24 inputs in both directions.
 * ------------------------------------------------------------------------------------------- */
/* ReinhardNormalizer.fit / get_mean_std (stainnorm.py:251-260, 337-367): mean and standard deviation of the float32 Lab channels
 * (L / 2.55, a - 128, b - 128) of one H*W*3 uint8 image on the device.  Host outputs; synchronises. */
int tiab_reinhard_fit(tiab_handle_t h, const uint8_t* target_dev, int H, int W, double means_host[3], double stds_host[3], void* stream);
/* ReinhardNormalizer.transform (stainnorm.py:262-284) over N <= 65535 images, in_dev / out_dev N*H*W*3 uint8 (must not overlap).  Optional device
 * outputs: src_stats_dev N*6 doubles {mean L, a, b, std L, a, b} of every source image; status_dev N ints — 1 when a channel of the
 * image is constant (the reference raises ZeroDivisionError at stainnorm.py:277-283); such an image is copied through unchanged. */
int tiab_reinhard_transform_batch(tiab_handle_t h, const uint8_t* in_dev, uint8_t* out_dev, int N, int H, int W, const double target_means[3],
                                  const double target_stds[3], double* src_stats_dev, int* status_dev, void* stream);
/* Path of tiab_reinhard_transform_batch on this handle: 0 = three launches (default); 1 = experimental fused kernel, one CTA per image,
 * for images of at most 196 608 bytes whose size is a multiple of 16 bytes (e.g. 256 x 256 x 3); other sizes keep path 0. */
int tiab_reinhard_set_path(tiab_handle_t h, int path);
/* cv2.cvtColor(x, COLOR_RGB2LAB) (inverse = 0) or COLOR_LAB2RGB (inverse = 1) for npix uint8 pixels on the device. */
int tiab_lab_convert(tiab_handle_t h, const uint8_t* in_dev, uint8_t* out_dev, long long npix, int inverse, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Convolution engine (patch CNNs).  Replaces the torch.nn.Conv2d / BatchNorm2d / ReLU library calls
 * inside tiatoolbox/models/architecture/hovernet.py:61-262,330-454, unet.py:24-418,
 * vanilla.py:256-333 (the per-batch `model(x)` of infer_batch: vanilla.py:251, unet.py:457,
 * hovernet.py:896).  All tensors are NHWC; activations fp16, accumulation fp32.
 * ------------------------------------------------------------------------------------------- */
#define TIAB_MAX_TAPS 49

/* One shifted-GEMM convolution:  D[p, co] = sum_t sum_ci X[p + tap_off[t], ci] * W[co, t, ci]
 * over the virtual output rows p in [0, M) of the input grid (Hin x Win per image); rows whose
 * (y, x) fall outside [y0,y1) x [x0,x1) are discarded.  Epilogue per surviving element:
 *   v = D + bias[co] (+ residual) ; relu? ; out32 / out16 <- v ; out16b <- relu(v*scale2 + shift2)
 * written at destination row (img*Hd + y + dy)*Wd + x + dx with channel offset coff*. */
typedef struct {
  const void* x;        /* fp16 [x_rows, cin], row pitch x_ld elements */
  long long x_rows;
  long long x_ld;
  int cin;
  int x_lo_off;         /* 0: fp16 operands.  >0: element offset of the residual plane x_lo = x - fp16(x); the GEMM
                           then runs the 3-term split A_hi*W_hi + A_lo*W_hi + A_hi*W_lo (w holds [W_hi | W_lo] per tap) */
  const void* w;        /* fp16 [N, taps*planes*ceil64(cin)] K-major (BN folded), row pitch w_ld elements */
  long long w_ld;
  int M, N, taps;
  int tap_off[TIAB_MAX_TAPS];
  int Hin, Win, y0, y1, x0, x1;
  int Hd, Wd, dy, dx;
  int Hr, Wr, ry, rx;
  void* out16;
  int ld16, coff16, lo16;     /* lo16: channel offset of the residual plane of out16 (0 = none) */
  void* out32;
  int ld32, coff32;
  void* out16b;
  int ld16b, coff16b, lo16b;
  const void* scale2;   /* fp32 [N] */
  const void* shift2;   /* fp32 [N] */
  const void* bias;     /* fp32 [N] or NULL */
  const void* res32;    /* fp32 residual or NULL */
  const void* res16;    /* fp16 residual or NULL */
  int ldres, coffres, reslo;
  int relu;
  /* A-operand transform (pre-activation BN + ReLU fused into this GEMM's operand path, hovernet.py:84-86, 170-172):
   * a = relu?(x * pre_scale[c] + pre_shift[c]) with x fp32 (x_is_f32) or fp16 hi(+lo at x_lo_off); `split` selects the
   * 3-term split-fp16 product in this mode (w then holds [W_hi | W_lo] per tap). */
  const void* pre_scale; /* fp32 [cin] or NULL */
  const void* pre_shift; /* fp32 [cin] or NULL */
  int pre_relu, x_is_f32, split;
  /* Grouped launch (several networks branches of identical shape in one grid, e.g. HoVer-Net's np / hv / tp decoders):
   * the M output rows are `groups` consecutive blocks of group_rows rows; group g uses weight rows [g*N, (g+1)*N) of w,
   * bias / scale2 / shift2 + g*N and pre_scale / pre_shift + g*cin.  x_grouped == 0: every group reads the same input rows
   * [0, group_rows) of x (a shared input); != 0: x holds groups*group_rows rows.  groups <= 1: plain launch. */
  int groups, group_rows, x_grouped;
  /* Strided convolutions without im2col.  The producer of the input stores it space-to-depth (s2d_c > 0 on ITS descriptor:
   * out16 pixel (y+dy, x+dx) goes to cell ((y+dy)/2, (x+dx)/2) of the Hd x Wd destination grid, channel block
   * ((y+dy)%2*2 + (x+dx)%2) * s2d_c); a stride-2 KxK convolution is then a stride-1 shifted GEMM over the cell grid whose tap
   * (kh, kw) has row offset (kh/2)*Wd + kw/2 and reads the channel block tap_coff[t] = ((kh%2)*2 + kw%2) * C of each row. */
  int tap_coff[TIAB_MAX_TAPS];
  int s2d_c;
  /* Fused pointwise head on a N == 64 convolution (HoVer-Net's u0: BN, ReLU, 1x1 conv + bias, hovernet.py:487-497):
   *   head_out[g][row, j] = head_b[g][j] + sum_c relu(D[row, c] * head_scale[g][c] + head_shift[g][c]) * head_w[g][j][c],  j < head_nout[g] <= 8
   * with row counted inside group g's destination grid; the 64-channel tensor itself is never stored.  head_w: fp32 [groups][8][64],
   * head_b: [groups][8], head_scale / head_shift: [groups][64]. */
  const void* head_w;
  const void* head_b;
  const void* head_scale;
  const void* head_shift;
  void* head_out[4];
  int head_nout[4];
  int head_ld[4];
  /* fp32 [groups*N] or NULL: the accumulator of output channel co is multiplied by out_scale[co] before the bias.  Lets the
   * packed weights be stored pre-scaled by a power of two per output channel, so that W_lo = W - fp16(W) stays inside fp16's
   * normal range whatever the magnitude of the checkpoint's weights. */
  const void* out_scale;
} tiab_conv_desc_t;

int tiab_conv_gemm(tiab_handle_t h, const tiab_conv_desc_t* d, void* stream);
/* explicit im2col (strided convs, 3-channel stems): out [N*Ho*Wo, ldout] fp16, k = (kh*KW + kw)*C + c, lo plane at +out_lo */
int tiab_im2col(tiab_handle_t h, const void* src, int src_is_u8, float u8_scale, int src_lo, int N, int H, int W, int C,
                int ldsrc, int KH, int KW, int stride, int pad_t, int pad_l, int Ho, int Wo, void* out_f16, int ldout,
                int Kpad, int out_lo, void* stream);
/* dst = relu?(src*scale + shift) -> fp16(+lo) (per-unit pre-activation of dense blocks, hovernet.py:84-86) */
int tiab_bn_relu(tiab_handle_t h, const void* src, int src_is_f32, int ldsrc, int coffsrc, int src_lo, const float* scale,
                 const float* shift, int relu, long long rows, int C, void* dst_f16, int lddst, int coffdst, int dst_lo,
                 void* stream);
/* dst = relu?(bn?(up2x(a[window]) + b[window])) (hovernet.py:440-451, unet.py:405-411); a is a Ha x Wa window at
 * (ay, ax) of an Hag x Wag grid */
int tiab_upsample_add(tiab_handle_t h, const void* a_f16, int Ha, int Wa, int Hag, int Wag, int ay, int ax, int lda, int a_lo,
                      const void* b_f16, int Hb, int Wb, int by, int bx, int ldb, int b_lo, int N, int C, const float* scale,
                      const float* shift, int relu, void* dst_f16, int Hd, int Wd, int dy, int dx, int ldd, int dst_lo,
                      void* stream);
int tiab_maxpool3x3s2(tiab_handle_t h, const void* src_f16, int N, int H, int W, int C, int lds, int src_lo, void* dst_f16,
                      int Hd, int Wd, int dy, int dx, int ldd, int dst_lo, void* stream);
/* out[p, j] = sum_c relu?(x[p,c]*scale[c]+shift[c]) * Wt[j,c] + bias[j], nout <= 16 (hovernet.py:487-497, unet.py:353) */
int tiab_pointwise_head(tiab_handle_t h, const void* x_f16, long long rows, int C, int ldx, int x_lo, const float* scale,
                        const float* shift, int relu, const float* Wt, const float* bias, int nout, float* out, int ldout,
                        int coffout, void* stream);
/* global average pool + Linear + softmax (vanilla.py:312-316) */
int tiab_gap_fc_softmax(tiab_handle_t h, const void* x_f16, int N, int HW, int C, int ldx, int x_lo, const float* Wt,
                        const float* bias, int ncls, float* feat_out, float* probs, void* stream);
/* HoVerNet.infer_batch head activations (hovernet.py:897-909) */
int tiab_hovernet_heads(tiab_handle_t h, const float* np_logits, const float* hv_logits, const float* tp_logits, int nt,
                        long long rows, float* np_out, float* hv_out, float* tp_out, void* stream);
/* UNetModel.infer_batch tail (unet.py:458-465): softmax -> bilinear x2 -> centre crop */
int tiab_softmax_up2_crop(tiab_handle_t h, const float* logits, int N, int H, int W, int C, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Whole networks from C.  Replaces ModelABC.infer_batch (tiatoolbox/models/models_abc.py:105-131;
 * HoVerNet.infer_batch hovernet.py:862-910, UNetModel.infer_batch unet.py:420-468, CNNModel.infer_batch vanilla.py:215-253)
 * for a binder without Python: `blob` is a serialised device plan (tiatoolbox_b200/models/export.py::export_plan — buffer table,
 * packed weights / folded BatchNorm constants, launch list with relocatable pointers; host memory, copied during the call).
 * tiab_net_create uploads it, replays it once and captures it into a CUDA graph; tiab_net_infer copies N <= batch uint8 NHWC
 * images (device pointer) into the plan's input buffer, launches the graph on `stream` and copies the first N rows of every
 * output head into out_heads[k] (device pointers, float32; NULL = skip that head).  tiab_net_info: info[0..4] = batch, in_h, in_w,
 * in_c, n_outputs, then per output {ndim, dims[4]}.
 */
typedef struct tiab_net* tiab_net_t;
int tiab_net_create(tiab_handle_t h, const void* blob, size_t blob_bytes, tiab_net_t* out);
int tiab_net_info(tiab_net_t net, int* info, int capacity);
int tiab_net_infer(tiab_net_t net, const uint8_t* in_nhwc_u8, int N, void* const* out_heads, void* stream);
void tiab_net_destroy(tiab_net_t net);

/* ---------------------------------------------------------------------------------------------
 * WSI stitching on the device.  Replaces the Python block loops of
 * tiatoolbox/models/engine/semantic_segmentor.py: merge_batch_to_canvas :1141-1183, merge_horizontal :1186-1263,
 * merge_vertical_chunkwise :1398-1534.  blocks_dev: N*bh*bw*C float32 patch predictions (device); locs_host: N*4 int32
 * [xs, ys, xe, ye] (host).  Sums are formed in the reference's order, so float32 results equal numpy's bit for bit.
 * ------------------------------------------------------------------------------------------- */
/* one patch row -> canvas (H*W*C float32, every block placed at row 0) + count (H*W uint8); all-zero blocks are skipped and not counted */
int tiab_merge_batch_to_canvas(tiab_handle_t h, const float* blocks_dev, int N, int bh, int bw, int C, const int* locs_host, int H, int W,
                               float* canvas_dev, uint8_t* count_dev, void* stream);
/* the whole merge of SemanticSegmentor.infer_wsi for in-memory patches in row-major grid order (a new patch row starts when ys changes):
 * row canvases of width canvas_width, seam folding of neighbouring rows, count normalisation (0 -> 1), clip to out_H x out_W. */
int tiab_stitch_patches(tiab_handle_t h, const float* blocks_dev, int N, int bh, int bw, int C, const int* locs_host, int canvas_width,
                        int out_H, int out_W, float* prob_dev, void* stream);

/* ---------------------------------------------------------------------------------------------
 * HoVer-Net post-processing on the device: HoVerNet._proc_np_hv + get_instance_info (without contours)
 * (tiatoolbox/models/architecture/hovernet.py:502-748).  np_map N*H*W*1, hv_map N*H*W*2, tp_map N*H*W*1 (or NULL),
 * all float32 device pointers as returned by infer_batch.  Outputs (device): inst_map N*H*W int32;
 * optional instance table (pass inst_count = NULL to skip): per image up to max_inst rows in ascending id order:
 * inst_id, inst_box [x0,y0,x1,y1) , inst_centroid [x,y] float64, inst_type, inst_prob float64.
 * ------------------------------------------------------------------------------------------- */
int tiab_hovernet_postproc(tiab_handle_t h, const float* np_map, const float* hv_map, const float* tp_map, int N, int H, int W,
                           int32_t* inst_map, int max_inst, int* inst_count, int32_t* inst_id, int32_t* inst_box,
                           double* inst_centroid, int32_t* inst_type, double* inst_prob, void* stream);
/* The same with HoVerNet._proc_np_hv's `scale_factor` (hovernet.py:506, 564-565): Sobel size int(20 * scale_factor + 1) and marker
 * size threshold ceil(10 * scale_factor^2) - 1.  Supported: 1.0 (HoVer-Net) and 0.5 (HoVerNetPlus.postproc, hovernetplus.py:305). */
int tiab_hovernet_postproc_scaled(tiab_handle_t h, const float* np_map, const float* hv_map, const float* tp_map, int N, int H, int W,
                                  double scale_factor, int32_t* inst_map, int max_inst, int* inst_count, int32_t* inst_id,
                                  int32_t* inst_box, double* inst_centroid, int32_t* inst_type, double* inst_prob, void* stream);
/* The contour part of get_instance_info (hovernet.py:685-693: cv2.findContours(crop, RETR_TREE, CHAIN_APPROX_SIMPLE)[0][0]) for every
 * row of the instance table written by tiab_hovernet_postproc: cont_xy N*capacity*2 int32 (x, y in image coordinates), the vertices of
 * table row (n, k) at [cont_off[n*max_inst+k], +cont_len[...]) of image n's slice; cont_total[n] = vertices image n needs (a slice that
 * would pass `capacity` is counted but not written).  All pointers are device pointers; asynchronous on `stream`. */
int tiab_hovernet_contours(tiab_handle_t h, const int32_t* inst_map, int N, int H, int W, int max_inst, const int* inst_count,
                           const int32_t* inst_id, const int32_t* inst_box, int capacity, int32_t* cont_off, int32_t* cont_len,
                           int32_t* cont_total, int32_t* cont_xy, void* stream);
/* Flood implementation of tiab_hovernet_postproc on this handle: 1 = binary heap, one thread per foreground component (default);
 * 2 = 8-ary heap with the total order (value, age, pixel index), one warp per component (csrc/watershed_math.cuh);
 * 3 = the same flood with the component's bounding box and heap staged in shared memory, persistent blocks over a work list;
 * 4 = work list + shared memory like 3, but the whole warp drives a 32-ary heap (children compared with redux.sync), the four neighbour
 *     tests run on four lanes. */
int tiab_hovernet_set_watershed(tiab_handle_t h, int impl);
int tiab_hovernet_get_watershed(tiab_handle_t h);   /* the implementation in use (1..4), negative on error */
/* host-side test hook of csrc/watershed_math.cuh: skimage.segmentation.watershed(dist, markers, mask=mask) on H*W host arrays (float64
 * image, int32 markers, uint8 mask) through the version-2 flood code. */
int tiab_dbg_watershed(const double* dist_host, const int32_t* markers_host, const uint8_t* mask_host, int H, int W, int32_t* out_host);
/* the work-list floods (impl 3: lane-0 8-ary heap in shared memory, 4: warp-cooperative 32-ary heap) on the DEVICE for arbitrary host
 * inputs (tests: adversarial tie cases on the GPU); whole mask = one work item */
int tiab_dbg_watershed_device(tiab_handle_t h, const double* dist_host, const int32_t* markers_host, const uint8_t* mask_host, int H, int W,
                              int impl, int32_t* out_host);
/* tiab_dbg_watershed through version 3's data flow on the host (per-component box-local copies, local raveled indices) */
int tiab_dbg_watershed_local(const double* dist_host, const int32_t* markers_host, const uint8_t* mask_host, int H, int W, int32_t* out_host);
int tiab_dbg_sobel_kernels(double deriv[21], double smooth[21]);
int tiab_dbg_sobel_kernels_k(int ksize, double* deriv, double* smooth);
/* host-side test hook of csrc/contour_math.cuh: the CHAIN_APPROX_SIMPLE outer border of label `id` (its component that contains the
 * label's first pixel in raster order) in an H*W int32 host image; returns the number of points (at most `cap` are written). */
int tiab_dbg_trace_contour(const int32_t* map_host, int H, int W, int32_t id, int* out_xy, int cap);
/* host-side test hooks of csrc/lab_math.cuh (no GPU): the device's per-pixel colour conversion and per-level Reinhard maps run on the CPU.
 * tiab_dbg_reinhard_maps returns 1 (and identity maps) when a channel is constant. */
int tiab_dbg_lab_convert(const uint8_t* in_host, uint8_t* out_host, long long npix, int inverse);
int tiab_dbg_reinhard_maps(const unsigned hist_host[768], double npix, const double target_means[3], const double target_stds[3],
                           double stats_out[6], uint8_t lut_out[768]);
int tiab_dbg_postproc_fetch(tiab_handle_t h, int which, void* host_out, void* stream);

/* Force one of the two device paths (testing / benchmarking): 0 = auto, 1 = fused one-CTA-per-patch
 * kernel (requires H*W*3 <= 196608 and H*W % 16 == 0), 2 = multi-pass general path. */
int tiab_stain_set_path(tiab_handle_t h, int path);
/* number of kernel launches issued by the last stain call on this handle */
int tiab_last_launch_count(tiab_handle_t h);

/* Host-side scalar pieces exported for the CPU unit tests (no GPU needed): the same
 * __host__ __device__ code the kernels' finalise steps run. */
int tiab_dbg_percentile_u8(const unsigned* hist256, double q, double* out);
int tiab_dbg_contrast_lut(const unsigned* hist256, unsigned n, uint8_t lut_out[256]);
int tiab_dbg_eigh3(const double cov[9], double evals[3], double evecs[9]);
int tiab_dbg_lasso2(const double gram[4], const double c[2], double reg, double n, double out[2]);
int tiab_dbg_macenko_from_angles(const double evec[6], double phi_min, double phi_max, double M_out[6]);
int tiab_dbg_tables(unsigned tabg[256], float od[256]);

#ifdef __cplusplus
}
#endif
#endif /* TIAB200_H */
