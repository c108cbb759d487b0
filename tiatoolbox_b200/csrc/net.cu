// net.cu — run a whole network from ONE C call (tiab_net_create / tiab_net_infer).
//
// Replaces, for a non-Python binder, ModelABC.infer_batch (tiatoolbox/models/models_abc.py:105-131; HoVerNet.infer_batch
// hovernet.py:862-910, UNetModel.infer_batch unet.py:420-468, CNNModel.infer_batch vanilla.py:215-253): the network is a serialised
// device plan (tiatoolbox_b200/models/export.py — buffer table, constants, launch list with relocatable pointers) that this file loads
// into HBM, replays once eagerly, captures into a CUDA graph and then launches per batch.  The launches themselves are the public
// per-layer entry points of this library (tiab_conv_gemm: tcgen05 shifted-GEMM convolution, ...), so the C path and the Python plan
// execute the same kernels with the same arguments.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/tiab200.h"
#include "ctx.h"

namespace {

struct NetArg {
  uint32_t type;   // 0 int, 1 double, 2 pointer (a = buffer, b = offset), 3 null, 4 conv descriptor
  uint64_t a, b;
};
struct NetOp {
  uint32_t func;
  std::vector<NetArg> args;
  tiab_conv_desc_t desc;
};
struct NetOut {
  uint32_t buf;
  uint64_t off;
  uint32_t ndim, dims[4], dtype;
  size_t row_bytes;   // bytes of one batch row
};

struct Reader {
  const uint8_t* p;
  size_t n, pos = 0;
  bool ok = true;
  template <class T> T get() {
    T v{};
    if (pos + sizeof(T) > n) {
      ok = false;
      return v;
    }
    memcpy(&v, p + pos, sizeof(T));
    pos += sizeof(T);
    return v;
  }
};

}  // namespace

struct tiab_net {
  tiab_ctx* h = nullptr;
  uint32_t batch = 0, in_h = 0, in_w = 0, in_c = 0, input_buf = 0;
  std::vector<void*> bufs;
  std::vector<size_t> sizes;
  std::vector<NetOp> ops;
  std::vector<NetOut> outs;
  cudaStream_t stream = nullptr;
  cudaGraphExec_t exec = nullptr;
  cudaGraph_t graph = nullptr;
};

namespace {

inline void* arg_ptr(const tiab_net* net, const NetArg& a) {
  return a.type == 2 ? (void*)((uint8_t*)net->bufs[a.a] + a.b) : nullptr;
}
inline long long arg_i(const NetArg& a) { return (long long)a.a; }
inline double arg_f(const NetArg& a) {
  double d;
  memcpy(&d, &a.a, 8);
  return d;
}

int run_op(tiab_net* net, const NetOp& op, void* s) {
  tiab_handle_t h = net->h;
  const std::vector<NetArg>& A = op.args;
  auto P = [&](int i) { return arg_ptr(net, A[i]); };
  auto I = [&](int i) { return (int)arg_i(A[i]); };
  auto L = [&](int i) { return arg_i(A[i]); };
  auto F = [&](int i) { return (float)arg_f(A[i]); };
  switch (op.func) {
    case 1:
      return tiab_conv_gemm(h, &op.desc, s);
    case 2:
      if (A.size() != 20) break;
      return tiab_im2col(h, P(0), I(1), F(2), I(3), I(4), I(5), I(6), I(7), I(8), I(9), I(10), I(11), I(12), I(13), I(14), I(15), P(16),
                         I(17), I(18), I(19), s);
    case 3:
      if (A.size() != 14) break;
      return tiab_bn_relu(h, P(0), I(1), I(2), I(3), I(4), (const float*)P(5), (const float*)P(6), I(7), L(8), I(9), P(10), I(11), I(12),
                          I(13), s);
    case 4:
      if (A.size() != 28) break;
      return tiab_upsample_add(h, P(0), I(1), I(2), I(3), I(4), I(5), I(6), I(7), I(8), P(9), I(10), I(11), I(12), I(13), I(14), I(15),
                               I(16), I(17), (const float*)P(18), (const float*)P(19), I(20), P(21), I(22), I(23), I(24), I(25), I(26),
                               I(27), s);
    case 5:
      if (A.size() != 14) break;
      return tiab_maxpool3x3s2(h, P(0), I(1), I(2), I(3), I(4), I(5), I(6), P(7), I(8), I(9), I(10), I(11), I(12), I(13), s);
    case 6:
      if (A.size() != 14) break;
      return tiab_pointwise_head(h, P(0), L(1), I(2), I(3), I(4), (const float*)P(5), (const float*)P(6), I(7), (const float*)P(8),
                                 (const float*)P(9), I(10), (float*)P(11), I(12), I(13), s);
    case 7:
      if (A.size() != 11) break;
      return tiab_gap_fc_softmax(h, P(0), I(1), I(2), I(3), I(4), I(5), (const float*)P(6), (const float*)P(7), I(8), (float*)P(9),
                                 (float*)P(10), s);
    case 8:
      if (A.size() != 8) break;
      return tiab_hovernet_heads(h, (const float*)P(0), (const float*)P(1), (const float*)P(2), I(3), L(4), (float*)P(5), (float*)P(6),
                                 (float*)P(7), s);
    case 9:
      if (A.size() != 6) break;
      return tiab_softmax_up2_crop(h, (const float*)P(0), I(1), I(2), I(3), I(4), (float*)P(5), s);
    default:
      break;
  }
  h->err = "tiab_net: malformed launch record (function " + std::to_string(op.func) + ", " + std::to_string(A.size()) + " arguments)";
  return TIAB_E_INVALID;
}

int run_all(tiab_net* net, void* s) {
  for (const NetOp& op : net->ops) {
    const int rc = run_op(net, op, s);
    if (rc != TIAB_OK) return rc;
  }
  return TIAB_OK;
}

}  // namespace

extern "C" {

void tiab_net_destroy(tiab_net_t net) {
  if (!net) return;
  TIAB_ON_DEVICE(net->h);
  if (net->stream) cudaStreamSynchronize(net->stream);
  if (net->exec) cudaGraphExecDestroy(net->exec);
  if (net->graph) cudaGraphDestroy(net->graph);
  for (void* p : net->bufs)
    if (p) cudaFree(p);
  if (net->stream) cudaStreamDestroy(net->stream);
  delete net;
}

int tiab_net_create(tiab_handle_t h, const void* blob, size_t blob_bytes, tiab_net_t* out) {
  if (!h || !blob || !out) return TIAB_E_INVALID;
  *out = nullptr;
  Reader r{(const uint8_t*)blob, blob_bytes};
  char magic[8];
  for (char& c : magic) c = (char)r.get<uint8_t>();
  if (!r.ok || memcmp(magic, "TIABPLN1", 8) != 0) {
    h->err = "tiab_net_create: not a tiab200 plan blob (magic TIABPLN1 expected)";
    return TIAB_E_INVALID;
  }
  TIAB_ON_DEVICE(h);
  tiab_net* net = new (std::nothrow) tiab_net();
  if (!net) return TIAB_E_NOMEM;
  net->h = h;
  const uint32_t n_buf = r.get<uint32_t>(), n_ops = r.get<uint32_t>();
  net->batch = r.get<uint32_t>(), net->in_h = r.get<uint32_t>(), net->in_w = r.get<uint32_t>(), net->in_c = r.get<uint32_t>();
  net->input_buf = r.get<uint32_t>();
  const uint32_t n_out = r.get<uint32_t>();
  auto fail = [&](const std::string& msg, int rc) {
    h->err = "tiab_net_create: " + msg;
    tiab_net_destroy(net);
    return rc;
  };
  if (!r.ok || n_buf == 0 || n_buf > (1u << 20) || n_ops > (1u << 20) || n_out == 0 || n_out > 8 || net->input_buf >= n_buf)
    return fail("corrupt header", TIAB_E_INVALID);
  for (uint32_t i = 0; i < n_out; ++i) {
    NetOut o{};
    o.buf = r.get<uint32_t>(), o.off = r.get<uint64_t>(), o.ndim = r.get<uint32_t>();
    for (uint32_t& d : o.dims) d = r.get<uint32_t>();
    o.dtype = r.get<uint32_t>();
    if (!r.ok || o.buf >= n_buf || o.ndim < 1 || o.ndim > 4 || o.dims[0] != net->batch) return fail("corrupt output table", TIAB_E_INVALID);
    o.row_bytes = 4;
    for (uint32_t d = 1; d < o.ndim; ++d) o.row_bytes *= o.dims[d];
    net->outs.push_back(o);
  }
  struct BufRec {
    uint64_t nbytes;
    uint32_t kind;
    uint64_t data_off;
  };
  std::vector<BufRec> recs(n_buf);
  for (BufRec& b : recs) b.nbytes = r.get<uint64_t>(), b.kind = r.get<uint32_t>(), b.data_off = r.get<uint64_t>();
  struct DescRef {
    size_t op;
    uint64_t off, len;
  };
  std::vector<DescRef> descs;
  net->ops.resize(n_ops);
  for (uint32_t i = 0; i < n_ops && r.ok; ++i) {
    NetOp& op = net->ops[i];
    op.func = r.get<uint32_t>();
    const uint32_t na = r.get<uint32_t>();
    if (na > 64) return fail("corrupt launch list", TIAB_E_INVALID);
    memset(&op.desc, 0, sizeof(op.desc));
    for (uint32_t k = 0; k < na; ++k) {
      NetArg a{};
      a.type = r.get<uint32_t>(), a.a = r.get<uint64_t>(), a.b = r.get<uint64_t>();
      if (a.type == 4)
        descs.push_back({i, a.a, a.b});
      else {
        if (a.type == 2 && (a.a >= n_buf || a.b >= recs[a.a].nbytes)) return fail("pointer outside its buffer", TIAB_E_INVALID);
        op.args.push_back(a);
      }
    }
  }
  if (!r.ok) return fail("truncated blob", TIAB_E_INVALID);
  r.pos = (r.pos + 15) & ~(size_t)15;
  const uint64_t data_len = r.get<uint64_t>();
  r.get<uint64_t>();
  const uint8_t* data = r.p + r.pos;
  if (!r.ok || r.pos + data_len > blob_bytes) return fail("truncated data section", TIAB_E_INVALID);
  // ---- buffers: scratch zero-initialised (padded borders of the activation buffers rely on it), constants uploaded ----
  net->bufs.assign(n_buf, nullptr);
  net->sizes.assign(n_buf, 0);
  for (uint32_t i = 0; i < n_buf; ++i) {
    const size_t nb = (size_t)recs[i].nbytes;
    net->sizes[i] = nb;
    if (cudaMalloc(&net->bufs[i], nb ? nb : 16) != cudaSuccess) return fail("out of device memory", TIAB_E_NOMEM);
    if (recs[i].kind == 1) {
      if (recs[i].data_off + nb > data_len) return fail("constant outside the data section", TIAB_E_INVALID);
      if (cudaMemcpy(net->bufs[i], data + recs[i].data_off, nb, cudaMemcpyHostToDevice) != cudaSuccess) return fail("upload failed", TIAB_E_CUDA);
    } else if (nb && cudaMemset(net->bufs[i], 0, nb) != cudaSuccess) {
      return fail("memset failed", TIAB_E_CUDA);
    }
  }
  // ---- conv descriptors: raw struct + pointer relocations ----
  for (const DescRef& d : descs) {
    if (d.len != sizeof(tiab_conv_desc_t) || d.off + d.len + 4 > data_len)
      return fail("conv descriptor size mismatch (plan exported by another library version)", TIAB_E_INVALID);
    NetOp& op = net->ops[d.op];
    memcpy(&op.desc, data + d.off, sizeof(tiab_conv_desc_t));
    uint32_t nrel;
    memcpy(&nrel, data + d.off + d.len, 4);
    const uint8_t* rp = data + d.off + d.len + 4;
    if (nrel > 64 || d.off + d.len + 4 + (uint64_t)nrel * 16 > data_len) return fail("corrupt relocation table", TIAB_E_INVALID);
    // every pointer field of the recorded descriptor held an address of the exporting process: clear them, then patch the relocated ones
    std::vector<uint32_t> fields;
    for (uint32_t k = 0; k < nrel; ++k) {
      uint32_t foff, buf;
      uint64_t off;
      memcpy(&foff, rp + 16 * k, 4), memcpy(&buf, rp + 16 * k + 4, 4), memcpy(&off, rp + 16 * k + 8, 8);
      if (foff + 8 > sizeof(tiab_conv_desc_t) || buf >= n_buf || off >= net->sizes[buf]) return fail("corrupt relocation", TIAB_E_INVALID);
      void* p = (uint8_t*)net->bufs[buf] + off;
      memcpy((uint8_t*)&op.desc + foff, &p, 8);
    }
  }
  if (cudaStreamCreateWithFlags(&net->stream, cudaStreamNonBlocking) != cudaSuccess) return fail("stream creation failed", TIAB_E_CUDA);
  // ---- one eager pass (function attributes, tensor-map entry point), then the capture ----
  int rc = run_all(net, net->stream);
  if (rc != TIAB_OK) {
    const std::string e = h->err;
    return fail("eager pass: " + e, rc);
  }
  if (cudaStreamSynchronize(net->stream) != cudaSuccess) return fail(std::string("eager pass: ") + cudaGetErrorString(cudaGetLastError()), TIAB_E_CUDA);
  if (cudaStreamBeginCapture(net->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) return fail("capture failed", TIAB_E_CUDA);
  rc = run_all(net, net->stream);
  cudaError_t ce = cudaStreamEndCapture(net->stream, &net->graph);
  if (rc != TIAB_OK || ce != cudaSuccess) return fail("graph capture failed", rc != TIAB_OK ? rc : TIAB_E_CUDA);
  if (cudaGraphInstantiate(&net->exec, net->graph, 0) != cudaSuccess) return fail("graph instantiation failed", TIAB_E_CUDA);
  *out = net;
  return TIAB_OK;
}

int tiab_net_info(tiab_net_t net, int* info, int capacity) {
  if (!net || !info || capacity < 5 + 5 * (int)net->outs.size()) return TIAB_E_INVALID;
  info[0] = (int)net->batch, info[1] = (int)net->in_h, info[2] = (int)net->in_w, info[3] = (int)net->in_c, info[4] = (int)net->outs.size();
  for (size_t k = 0; k < net->outs.size(); ++k) {
    info[5 + 5 * k] = (int)net->outs[k].ndim;
    for (int d = 0; d < 4; ++d) info[5 + 5 * k + 1 + d] = (int)net->outs[k].dims[d];
  }
  return TIAB_OK;
}

int tiab_net_infer(tiab_net_t net, const uint8_t* in_nhwc_u8, int N, void* const* out_heads, void* stream) {
  if (!net) return TIAB_E_INVALID;
  tiab_handle_t h = net->h;
  if (!in_nhwc_u8 || !out_heads || N <= 0 || N > (int)net->batch) {
    h->err = "tiab_net_infer: invalid argument (1 <= N <= plan batch, device pointers required)";
    return TIAB_E_INVALID;
  }
  TIAB_ON_DEVICE(h);
  cudaStream_t s = (cudaStream_t)stream;
  const size_t in_row = (size_t)net->in_h * net->in_w * net->in_c;
  TIAB_CUDA_OK(h, cudaMemcpyAsync(net->bufs[net->input_buf], in_nhwc_u8, in_row * N, cudaMemcpyDeviceToDevice, s));
  TIAB_CUDA_OK(h, cudaGraphLaunch(net->exec, s));
  for (size_t k = 0; k < net->outs.size(); ++k) {
    if (!out_heads[k]) continue;
    const NetOut& o = net->outs[k];
    TIAB_CUDA_OK(h, cudaMemcpyAsync(out_heads[k], (uint8_t*)net->bufs[o.buf] + o.off, o.row_bytes * N, cudaMemcpyDeviceToDevice, s));
  }
  h->last_launches = (int)net->ops.size();
  return TIAB_OK;
}

}  // extern "C"
