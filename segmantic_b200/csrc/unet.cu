// sgm_unet handle: weight packing, the per-window-batch layer program, and the C ABI for
// Net.forward (seg/monai_unet.py:221-222) and the sliding-window inferer (seg/monai_unet.py:637-665).
#include "common.cuh"
#include "conv_tc.cuh"

#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <vector>

namespace sgm {

static std::recursive_mutex g_mutex;
void lock_global() { g_mutex.lock(); }
void unlock_global() { g_mutex.unlock(); }

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int launch_finalize(const float* acc, int channels, const sgm_sw_cfg* cfg, const int* starts_dev,
                    const float* imap_dev[3], float* logits, uint8_t* labels, float* probs,
                    cudaStream_t st);
int launch_gather_blend(const float* wl, int channels, const sgm_sw_cfg* cfg, const int* starts_dev,
                        const float* imap_dev[3], float* logits, uint8_t* labels, float* probs, cudaStream_t st);

struct PackedConv {
  int kind = SGM_KIND_IDENTITY;
  int cin = 0, cout = 0, cgin = 0, cgout = 0;
  int k[3] = {1, 1, 1}, s[3] = {1, 1, 1}, pad[3] = {0, 0, 0};
  int act = 0;
  float alpha = 0.f;
  float* w32 = nullptr;   // CUDA-core family packing
  float* bias = nullptr;  // [cgout*8] (padded to the cout tile)
  tc::TcConv* tc = nullptr;        // tcgen05 packing of this conv (bf16 precision)
  tc::TcConv* tc_fused = nullptr;  // unit0 of a strided down block fused with its residual-branch conv
};

}  // namespace sgm

struct sgm_unet {
  int spatial_dims = 3, cin = 1, cout = 1, n_levels = 0;
  int channels[SGM_MAX_LEVELS] = {0}, strides[SGM_MAX_LEVELS] = {0};
  int precision = SGM_PRECISION_FP32;
  std::vector<sgm::PackedConv> convs;
  int64_t last_launches = 0;
  // tcgen05 dispatch mask (env SGM_TC, default all): 1 stride-1 convs, 2 strided down convs (fused
  // with the residual branch), 4 transposed convs, 8 head (conv + blend epilogue)
  float* stem_w = nullptr;     // fused stem: [coblk][tap][ci][STEM_CO] + bias
  float* stem_bias = nullptr;
  void* stem_tc_w = nullptr;   // tensor-core stem packing (bf16 precision, eligible first blocks only)
  float* stem_tc_b = nullptr;
  int stem_tc_kp = 0;
  int tc_mask = 15;
  int* err_dev = nullptr;  // device flag raised by a tcgen05 pipeline timeout
  // optional per-convolution CUDA-event timing (sgm_unet_set_profiling)
  struct ProfEv {
    cudaEvent_t a, b;
    int conv;
  };
  bool profiling = false;
  std::vector<ProfEv> prof_pending;
  std::vector<cudaEvent_t> ev_pool;
  std::vector<double> prof_ms;
  std::vector<int64_t> prof_n;
  cudaEvent_t ev_get() {
    cudaEvent_t e = nullptr;
    if (!ev_pool.empty()) {
      e = ev_pool.back();
      ev_pool.pop_back();
    } else {
      cudaEventCreate(&e);
    }
    return e;
  }
};

namespace sgm {

namespace {

// round-to-nearest-even fp32 -> bf16 -> fp32 (host)
inline float bf16_round(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7f800000u) != 0x7f800000u) u += 0x7fffu + ((u >> 16) & 1u);
  u &= 0xffff0000u;
  memcpy(&f, &u, 4);
  return f;
}

// ---- CUDA-core packing: [coblk][cg][tap][ci 8][co CO_T], zero padded (bf16 path: bf16-rounded values)
int pack_fp32(const sgm_conv_desc& d, PackedConv& pc, bool round_bf16) {
  const bool tr = pc.kind == SGM_KIND_CONV_TRANSPOSE;
  const int co_t = fp32_conv_cout_tile(tr && d.stride == 2);  // stride-1 convT runs as a flipped conv
  const int ntaps = pc.k[0] * pc.k[1] * pc.k[2];
  const int n_coblk = ceil_div(pc.cgout * 8, co_t);
  const size_t nw = (size_t)n_coblk * pc.cgin * ntaps * 8 * co_t;
  std::vector<float> w(nw, 0.f), b((size_t)n_coblk * co_t, 0.f);
  // source taps: d.weight is [O][I][kk] (conv) or [I][O][kk] (convT) with kk = kernel^spatial_dims
  const bool flip = tr && d.stride == 1;  // stride-1 transposed conv == conv with flipped kernel
  for (int co = 0; co < pc.cout; ++co) {
    b[co] = d.bias[co];
    for (int ci = 0; ci < pc.cin; ++ci)
      for (int t = 0; t < ntaps; ++t) {
        const int st = flip ? (ntaps - 1 - t) : t;
        const float v = tr ? d.weight[((size_t)ci * pc.cout + co) * ntaps + st]
                           : d.weight[((size_t)co * pc.cin + ci) * ntaps + st];
        const size_t dst = ((((size_t)(co / co_t) * pc.cgin + ci / 8) * ntaps + t) * 8 + ci % 8) * co_t + co % co_t;
        w[dst] = round_bf16 ? bf16_round(v) : v;
      }
  }
  if (cudaMalloc(&pc.w32, nw * sizeof(float)) != cudaSuccess ||
      cudaMalloc(&pc.bias, b.size() * sizeof(float)) != cudaSuccess) {
    set_error("cudaMalloc of packed weights failed: %s", cudaGetErrorString(cudaGetLastError()));
    return SGM_ERR_CUDA;
  }
  SGM_CUDA_CHECK(cudaMemcpy(pc.w32, w.data(), nw * sizeof(float), cudaMemcpyHostToDevice));
  SGM_CUDA_CHECK(cudaMemcpy(pc.bias, b.data(), b.size() * sizeof(float), cudaMemcpyHostToDevice));
  return SGM_OK;
}

// ---- fused stem packing: [coblk][tap][ci][STEM_CO]; fused channel f < cgA*8 is unit0's (padded),
// the rest the residual-branch conv's.  A k=1 residual conv (stride-1 stem) sits on the centre tap.
int pack_stem(const sgm_conv_desc& u0, const sgm_conv_desc& rs, const PackedConv& p0, const PackedConv& pr,
              bool round_bf16, float** w_dev, float** b_dev) {
  if (rs.kind != SGM_KIND_CONV) return SGM_OK;  // identity residual: no fused stem (rejected at run time)
  const int co_t = stem_cout_tile();
  const int ntaps = p0.k[0] * p0.k[1] * p0.k[2];
  const int nf = (p0.cgout + pr.cgout) * 8;
  const int ncoblk = ceil_div(nf, co_t);
  std::vector<float> w((size_t)ncoblk * ntaps * u0.cin * co_t, 0.f), b((size_t)ncoblk * co_t, 0.f);
  const int rtaps = pr.k[0] * pr.k[1] * pr.k[2];
  for (int f = 0; f < nf; ++f) {
    const bool isA = f < p0.cgout * 8;
    const sgm_conv_desc& src = isA ? u0 : rs;
    const int co = isA ? f : f - p0.cgout * 8;
    if (co >= src.cout) continue;
    b[f] = src.bias[co];
    for (int ci = 0; ci < u0.cin; ++ci)
      for (int t = 0; t < ntaps; ++t) {
        float v;
        if (isA || rtaps == ntaps) {
          v = src.weight[((size_t)co * src.cin + ci) * ntaps + t];
        } else {  // k=1 residual: centre tap only
          v = (t == ntaps / 2) ? src.weight[(size_t)co * src.cin + ci] : 0.f;
        }
        w[(((size_t)(f / co_t) * ntaps + t) * u0.cin + ci) * co_t + f % co_t] = round_bf16 ? bf16_round(v) : v;
      }
  }
  if (cudaMalloc(w_dev, w.size() * 4) != cudaSuccess || cudaMalloc(b_dev, b.size() * 4) != cudaSuccess) {
    set_error("cudaMalloc of stem weights failed");
    return SGM_ERR_CUDA;
  }
  SGM_CUDA_CHECK(cudaMemcpy(*w_dev, w.data(), w.size() * 4, cudaMemcpyHostToDevice));
  SGM_CUDA_CHECK(cudaMemcpy(*b_dev, b.data(), b.size() * 4, cudaMemcpyHostToDevice));
  return SGM_OK;
}

struct Tensor {
  void* p = nullptr;
  int cg = 0;
  int d[3] = {0, 0, 0};
  long long vox() const { return (long long)d[0] * d[1] * d[2]; }
};

struct Bump {
  char* base;
  int64_t size, off = 0;
  bool dry;
  void* take(int64_t bytes) {
    off = (off + 255) & ~int64_t(255);
    void* p = dry ? nullptr : base + off;
    off += bytes;
    return p;
  }
};

struct HeadTarget {
  int kind;  // OUT_PLANAR or OUT_BLEND
  float* out;
  long long cstride, nstride;
  int ad0, ad1, ad2;
  const int* wo_host;  // [n][3] blend origins (host)
  const float* imap[3];
  float floor;
  int weighted;  // OUT_PLANAR: store logits * importance map (deferred blend)
};

inline int out_dim(int i, int k, int s, int pad) { return (i + 2 * pad - k) / s + 1; }

// Runs the whole network on `n` windows.  Input: planar volume + device window origins.
int run_network(sgm_unet* net, const float* vol, long long vol_cstride, int vd1, int vd2,
                const int* win_origin_dev, int n, const int roi[3], Bump& ws, const HeadTarget& head,
                cudaStream_t st, bool dry) {
  const int L = net->n_levels;
  const size_t esz = net->precision == SGM_PRECISION_BF16 ? 2 : 4;
  const bool bf16 = net->precision == SGM_PRECISION_BF16;
  struct Prof {  // records a CUDA-event pair around one launch when profiling is on
    sgm_unet* net;
    cudaStream_t st;
    cudaEvent_t b = nullptr;
    Prof(sgm_unet* n, cudaStream_t s, const PackedConv* pc, bool dry) : net(n), st(s) {
      if (dry || !n->profiling) return;
      sgm_unet::ProfEv e;
      e.a = n->ev_get(), e.b = n->ev_get(), e.conv = (int)(pc - n->convs.data());
      cudaEventRecord(e.a, s);
      b = e.b;
      n->prof_pending.push_back(e);
    }
    ~Prof() {
      if (b) cudaEventRecord(b, st);
    }
  };
  auto alloc = [&](int cg, const int d[3]) {
    Tensor t;
    t.cg = cg;
    t.d[0] = d[0], t.d[1] = d[1], t.d[2] = d[2];
    t.p = ws.take((int64_t)n * cg * t.vox() * 8 * esz);
    return t;
  };
  auto base_args = [&](const PackedConv& pc, const Tensor& in0, const Tensor* in1, const int od[3]) {
    ConvArgs a;
    memset(&a, 0, sizeof(a));
    a.in0 = in0.p, a.cg0 = in0.cg;
    a.in1 = in1 ? in1->p : nullptr, a.cg1 = in1 ? in1->cg : 0;
    a.cin_real = pc.cin;
    a.n = n;
    for (int i = 0; i < 3; ++i) {
      a.id[i] = in0.d[i], a.od[i] = od[i];
      a.k[i] = pc.k[i], a.s[i] = pc.s[i], a.pad[i] = pc.pad[i];
    }
    a.w = pc.w32, a.bias = pc.bias, a.cout_groups = pc.cgout;
    a.act = pc.act, a.alpha = pc.alpha;
    a.c_real = pc.cout;
    return a;
  };
  auto tc_bit = [&](const PackedConv& pc) {
    if (!pc.tc) return 0;
    return pc.tc->mode == tc::MODE_S1 ? 1 : (pc.tc->mode == tc::MODE_S2 ? 2 : 4);
  };
  auto conv_dims = [&](const PackedConv& pc, const Tensor& in0, int od[3]) {
    if (pc.kind == SGM_KIND_CONV_TRANSPOSE && pc.s[1] == 2) {
      for (int i = 0; i < 3; ++i) od[i] = in0.d[i] * pc.s[i];
    } else {
      for (int i = 0; i < 3; ++i) od[i] = out_dim(in0.d[i], pc.k[i], pc.s[i], pc.pad[i]);
    }
  };
  auto tc_io = [&](const Tensor& in0, const Tensor* in1, const int od[3]) {
    tc::TcIO io;
    io.in0 = in0.p, io.cg0 = in0.cg;
    io.in1 = in1 ? in1->p : nullptr, io.cg1 = in1 ? in1->cg : 0;
    io.n = n;
    for (int i = 0; i < 3; ++i) io.id[i] = in0.d[i], io.od[i] = od[i];
    return io;
  };
  auto conv = [&](const PackedConv& pc, const Tensor& in0, const Tensor* in1, Tensor& out,
                  const Tensor* res) -> int {
    int od[3];
    conv_dims(pc, in0, od);
    out = alloc(pc.cgout, od);
    if (dry) return SGM_OK;
    SGM_REQUIRE(in0.cg + (in1 ? in1->cg : 0) == pc.cgin, SGM_ERR_INVALID, "channel-group mismatch");
    net->last_launches++;
    Prof prof(net, st, &pc, dry);
    if (bf16 && (net->tc_mask & tc_bit(pc))) {
      tc::TcIO io = tc_io(in0, in1, od);
      io.outA = out.p, io.cgA = out.cg, io.res = res ? res->p : nullptr;
      return tc::tc_launch(*pc.tc, io, net->err_dev, st);
    }
    ConvArgs a = base_args(pc, in0, in1, od);
    a.out = out.p;
    a.res = res ? res->p : nullptr;
    if (pc.kind == SGM_KIND_CONV_TRANSPOSE && pc.s[1] == 2) return launch_convT_fp32(a, bf16, st);
    return launch_conv_fp32(a, bf16, false, OUT_CG8, st);
  };
  // unit0 + residual-branch conv of a strided down block: same input, one fused tcgen05 launch
  auto conv_pair = [&](const PackedConv& u0, const PackedConv& rs, const Tensor& in0, Tensor& t, Tensor& r) -> int {
    if (bf16 && u0.tc_fused && (net->tc_mask & 2)) {
      int od[3];
      conv_dims(u0, in0, od);
      t = alloc(u0.cgout, od);
      r = alloc(rs.cgout, od);
      if (dry) return SGM_OK;
      net->last_launches++;
      Prof prof(net, st, &u0, dry);
      tc::TcIO io = tc_io(in0, nullptr, od);
      io.outA = t.p, io.cgA = t.cg, io.outB = r.p, io.cgB = r.cg;
      return tc::tc_launch(*u0.tc_fused, io, net->err_dev, st);
    }
    int rc = conv(u0, in0, nullptr, t, nullptr);
    if (rc) return rc;
    if (rs.kind == SGM_KIND_IDENTITY) {
      r = in0;
      return SGM_OK;
    }
    return conv(rs, in0, nullptr, r, nullptr);
  };

  // level dims
  std::vector<PackedConv>& cv = net->convs;
  Tensor cur;  // current CG8 activation (invalid for the planar network input)
  std::vector<Tensor> skips(L);
  int idx = 0;
  for (int i = 0; i < L; ++i) {
    const PackedConv& u0 = cv[idx], &u1 = cv[idx + 1], &rs = cv[idx + 2];
    idx += 3;
    Tensor t, r, x;
    if (i == 0) {
      // stem: reads windows straight from the planar volume
      int od[3];
      for (int a = 0; a < 3; ++a) od[a] = out_dim(roi[a], u0.k[a], u0.s[a], u0.pad[a]);
      SGM_REQUIRE(rs.kind == SGM_KIND_CONV, SGM_ERR_UNSUPPORTED,
                  "first down block with identity residual (stride 1 and num_channels == channels[0])");
      t = alloc(u0.cgout, od);
      r = alloc(rs.cgout, od);
      if (!dry) {
        ConvArgs a;
        memset(&a, 0, sizeof(a));
        a.in0 = vol, a.cin_real = u0.cin, a.n = n;
        for (int q = 0; q < 3; ++q) {
          a.id[q] = roi[q], a.od[q] = od[q];
          a.k[q] = u0.k[q], a.s[q] = u0.s[q], a.pad[q] = u0.pad[q];
        }
        a.w = net->stem_w, a.bias = net->stem_bias, a.act = u0.act, a.alpha = u0.alpha;
        a.vol_cstride = vol_cstride, a.vd1 = vd1, a.vd2 = vd2, a.win_origin = win_origin_dev;
        net->last_launches++;
        Prof prof(net, st, &u0, dry);
        int rc = (bf16 && net->stem_tc_w)
                     ? launch_stem_tc(a, net->stem_tc_w, net->stem_tc_b, net->stem_tc_kp, t.cg, r.cg, t.p, r.p, net->err_dev, st)
                     : launch_stem(a, t.cg, r.cg, t.p, r.p, bf16, st);
        if (rc) return rc;
      }
    } else {
      int rc = conv_pair(u0, rs, cur, t, r);
      if (rc) return rc;
    }
    int rc = conv(u1, t, nullptr, x, &r);
    if (rc) return rc;
    skips[i] = x;
    cur = x;
  }
  // bottom
  Tensor sub;
  {
    const PackedConv& u0 = cv[idx], &u1 = cv[idx + 1], &rs = cv[idx + 2];
    idx += 3;
    Tensor t, r;
    int rc = conv(u0, cur, nullptr, t, nullptr);
    if (rc) return rc;
    if (rs.kind == SGM_KIND_IDENTITY) {
      r = cur;
    } else {
      rc = conv(rs, cur, nullptr, r, nullptr);
      if (rc) return rc;
    }
    rc = conv(u1, t, nullptr, sub, &r);
    if (rc) return rc;
  }
  // up path
  for (int i = L - 1; i >= 0; --i) {
    const PackedConv& ct = cv[idx], &ru = cv[idx + 1];
    idx += 2;
    Tensor u;
    int rc = conv(ct, skips[i], &sub, u, nullptr);
    if (rc) return rc;
    if (i > 0) {
      for (int a = 0; a < 3; ++a)
        SGM_REQUIRE(u.d[a] == skips[i - 1].d[a], SGM_ERR_INVALID,
                    "roi is not compatible with the strides (skip/up shape mismatch at level %d)", i);
      Tensor o;
      rc = conv(ru, u, nullptr, o, &u);
      if (rc) return rc;
      sub = o;
    } else {
      for (int a = 0; a < 3; ++a)
        SGM_REQUIRE(u.d[a] == roi[a], SGM_ERR_INVALID,
                    "roi %d along axis %d is not restored by the up path (must be divisible by the "
                    "product of strides)", roi[a], a);
      if (dry) break;
      // head: conv C->C (conv only) + identity residual -> planar logits / blended accumulator
      const long long uvox = u.vox();
      const bool head_tc = bf16 && ru.tc && (net->tc_mask & 8);
      ConvArgs a = base_args(ru, u, nullptr, u.d);
      a.pl_out = head.out, a.pl_cstride = head.cstride, a.pl_nstride = head.nstride;
      a.ad0 = head.ad0, a.ad1 = head.ad1, a.ad2 = head.ad2;
      a.imap0 = head.imap[0], a.imap1 = head.imap[1], a.imap2 = head.imap[2];
      a.imap_floor = head.floor;
      a.pl_weighted = head.weighted;
      tc::TcIO io = tc_io(u, nullptr, u.d);
      io.pl_weighted = head.weighted;
      io.cgA = u.cg;
      io.pl_out = head.out, io.pl_cstride = head.cstride, io.pl_nstride = head.nstride;
      io.ad0 = head.ad0, io.ad1 = head.ad1, io.ad2 = head.ad2;
      for (int q = 0; q < 3; ++q) io.imap[q] = head.imap[q];
      io.imap_floor = head.floor;
      if (head.kind == OUT_PLANAR) {
        net->last_launches++;
        Prof prof(net, st, &ru, dry);
        if (head_tc) {
          io.res = u.p, io.out_kind = OUT_PLANAR;
          rc = tc::tc_launch(*ru.tc, io, net->err_dev, st);
        } else {
          a.res = u.p;
          rc = launch_conv_fp32(a, bf16, false, OUT_PLANAR, st);
        }
        if (rc) return rc;
      } else {
        for (int w = 0; w < n; ++w) {  // one launch per window: plain RMW, MONAI's window order
          const char* uw = reinterpret_cast<const char*>(u.p) + (size_t)w * u.cg * uvox * 8 * esz;
          net->last_launches++;
          Prof prof(net, st, &ru, dry);
          if (head_tc) {
            tc::TcIO b = io;
            b.n = 1, b.in0 = uw, b.res = uw, b.out_kind = OUT_BLEND;
            for (int q = 0; q < 3; ++q) b.wo[q] = head.wo_host[w * 3 + q];
            rc = tc::tc_launch(*ru.tc, b, net->err_dev, st);
          } else {
            ConvArgs b = a;
            b.n = 1, b.in0 = uw, b.res = uw;
            for (int q = 0; q < 3; ++q) b.wo[q] = head.wo_host[w * 3 + q];
            rc = launch_conv_fp32(b, bf16, false, OUT_BLEND, st);
          }
          if (rc) return rc;
        }
      }
    }
  }
  return SGM_OK;
}

int check_roi(const sgm_unet* net, const int32_t roi[3]) {
  SGM_REQUIRE(roi && roi[0] > 0 && roi[1] > 0 && roi[2] > 0, SGM_ERR_INVALID, "bad roi");
  SGM_REQUIRE(net->spatial_dims == 3 || roi[0] == 1, SGM_ERR_INVALID,
              "2-D networks take roi = {1, h, w}");
  return SGM_OK;
}

}  // namespace
}  // namespace sgm

using namespace sgm;

extern "C" const char* sgm_last_error(void) { return g_err; }
extern "C" int32_t sgm_version(void) { return 100; }

extern "C" void sgm_unet_destroy(sgm_unet* net) {
  if (!net) return;
  for (auto& c : net->convs) {
    if (c.w32) cudaFree(c.w32);
    if (c.bias) cudaFree(c.bias);
    tc::tc_free(c.tc);
    tc::tc_free(c.tc_fused);
  }
  if (net->err_dev) cudaFree(net->err_dev);
  if (net->stem_w) cudaFree(net->stem_w);
  if (net->stem_bias) cudaFree(net->stem_bias);
  if (net->stem_tc_w) cudaFree(net->stem_tc_w);
  if (net->stem_tc_b) cudaFree(net->stem_tc_b);
  for (auto& e : net->prof_pending) cudaEventDestroy(e.a), cudaEventDestroy(e.b);
  for (auto& e : net->ev_pool) cudaEventDestroy(e);
  delete net;
}

extern "C" int32_t sgm_unet_create(const sgm_unet_desc* d, sgm_unet** out) {
  SGM_REQUIRE(d && out, SGM_ERR_INVALID, "sgm_unet_create: null argument");
  *out = nullptr;
  SGM_REQUIRE(d->spatial_dims == 2 || d->spatial_dims == 3, SGM_ERR_INVALID, "spatial_dims must be 2 or 3");
  SGM_REQUIRE(d->n_levels >= 1 && d->n_levels < SGM_MAX_LEVELS, SGM_ERR_INVALID, "bad n_levels");
  SGM_REQUIRE(d->n_convs == 3 * d->n_levels + 3 + 2 * d->n_levels, SGM_ERR_INVALID,
              "expected %d convolutions in canonical order, got %d", 5 * d->n_levels + 3, d->n_convs);
  SGM_REQUIRE(d->precision == SGM_PRECISION_FP32 || d->precision == SGM_PRECISION_BF16, SGM_ERR_INVALID,
              "bad precision");
  SGM_REQUIRE(d->in_channels >= 1 && d->in_channels <= 8, SGM_ERR_UNSUPPORTED,
              "num_channels must be in 1..8, got %d", d->in_channels);
  SGM_REQUIRE(d->out_channels >= 1 && d->out_channels <= 64, SGM_ERR_UNSUPPORTED,
              "num_classes must be in 1..64, got %d", d->out_channels);
  int dev_count = 0;
  if (cudaGetDeviceCount(&dev_count) != cudaSuccess || dev_count == 0) {
    set_error("no CUDA device: segmantic_b200 has no CPU fallback");
    return SGM_ERR_CUDA;
  }
  for (int i = 0; i <= d->n_levels; ++i)
    SGM_REQUIRE(d->channels[i] > 0 && d->channels[i] % 8 == 0, SGM_ERR_UNSUPPORTED,
                "channels must be positive multiples of 8, got %d", d->channels[i]);
  sgm_unet* net = new sgm_unet();
  net->spatial_dims = d->spatial_dims, net->cin = d->in_channels, net->cout = d->out_channels;
  net->n_levels = d->n_levels, net->precision = d->precision;
  for (int i = 0; i <= d->n_levels; ++i) {
    net->channels[i] = d->channels[i];
    if (i < d->n_levels) net->strides[i] = d->strides[i];
  }
  const bool bf16 = d->precision == SGM_PRECISION_BF16;
  if (bf16) {
    for (int i = 0; i <= d->n_levels; ++i)
      if (d->channels[i] % 16 != 0) {
        set_error("precision bf16 needs channels that are multiples of 16, got %d", d->channels[i]);
        sgm_unet_destroy(net);
        return SGM_ERR_UNSUPPORTED;
      }
    if (const char* env = getenv("SGM_TC")) net->tc_mask = atoi(env);
    if (cudaMalloc(&net->err_dev, sizeof(int)) != cudaSuccess || cudaMemset(net->err_dev, 0, sizeof(int)) != cudaSuccess) {
      set_error("cudaMalloc(error flag) failed");
      sgm_unet_destroy(net);
      return SGM_ERR_CUDA;
    }
  }
  net->convs.resize(d->n_convs);
  for (int i = 0; i < d->n_convs; ++i) {
    const sgm_conv_desc& c = d->convs[i];
    PackedConv& pc = net->convs[i];
    pc.kind = c.kind;
    pc.cin = c.cin, pc.cout = c.cout;
    pc.cgin = ceil_div(c.cin, 8), pc.cgout = ceil_div(c.cout, 8);
    const bool stem = (i == 0 || i == 2);  // read the planar fp32 volume, never CG8
    if (bf16) {                            // bf16 CG8 tensors carry channels padded to 16 (even groups)
      if (!stem) pc.cgin = ceil_div(c.cin, 16) * 2;
      pc.cgout = ceil_div(c.cout, 16) * 2;
    }
    pc.act = c.has_act, pc.alpha = c.alpha;
    if (c.kind == SGM_KIND_IDENTITY) continue;
    if (!(c.kernel == 1 || c.kernel == 3) || !(c.stride == 1 || c.stride == 2) || !c.weight || !c.bias) {
      set_error("conv %d: unsupported kernel/stride or null weights", i);
      sgm_unet_destroy(net);
      return SGM_ERR_UNSUPPORTED;
    }
    for (int a = 0; a < 3; ++a) {
      const bool flat = d->spatial_dims == 2 && a == 0;
      pc.k[a] = flat ? 1 : c.kernel;
      pc.s[a] = flat ? 1 : c.stride;
      pc.pad[a] = flat ? 0 : c.kernel / 2;
    }
    int rc = pack_fp32(c, pc, bf16);
    if (!rc && bf16 && !stem && tc::tc_supported(c)) rc = tc::tc_pack(&c, nullptr, d->spatial_dims, &pc.tc);
    // strided down block (levels >= 1): unit0 (index 3l) fused with its residual conv (index 3l+2)
    if (!rc && bf16 && i >= 3 && i < 3 * d->n_levels && i % 3 == 0 && c.stride == 2 &&
        d->convs[i + 2].kind == SGM_KIND_CONV && d->convs[i + 2].kernel == c.kernel && tc::tc_supported(c))
      rc = tc::tc_pack(&c, &d->convs[i + 2], d->spatial_dims, &pc.tc_fused);
    if (rc) {
      sgm_unet_destroy(net);
      return rc;
    }
  }
  {
    int rc = pack_stem(d->convs[0], d->convs[2], net->convs[0], net->convs[2], bf16, &net->stem_w, &net->stem_bias);
    if (!rc && bf16)
      rc = stem_tc_pack(d->convs[0], d->convs[2], d->spatial_dims, &net->stem_tc_w, &net->stem_tc_b, &net->stem_tc_kp);
    if (rc) {
      sgm_unet_destroy(net);
      return rc;
    }
  }
  *out = net;
  return SGM_OK;
}

extern "C" int64_t sgm_unet_last_launch_count(const sgm_unet* net) { return net ? net->last_launches : 0; }

extern "C" int64_t sgm_unet_workspace_bytes(const sgm_unet* net, const int32_t roi[3], int32_t batch) {
  if (!net || check_roi(net, roi) || batch < 1) return SGM_ERR_INVALID;
  Bump ws{nullptr, 0, 0, true};
  ws.take((int64_t)batch * 3 * sizeof(int));
  HeadTarget head;
  memset(&head, 0, sizeof(head));
  int rc = run_network(const_cast<sgm_unet*>(net), nullptr, 0, 0, 0, nullptr, batch, roi, ws, head, 0, true);
  if (rc) return rc;
  return ws.off + 256;
}

extern "C" int32_t sgm_unet_forward(sgm_unet* net, const float* x_dev, float* logits_dev, int32_t batch,
                                    const int32_t roi[3], void* workspace_dev, int64_t workspace_bytes,
                                    void* stream) {
  SGM_REQUIRE(net && x_dev && logits_dev && workspace_dev && batch >= 1, SGM_ERR_INVALID,
              "sgm_unet_forward: bad argument");
  int rc = check_roi(net, roi);
  if (rc) return rc;
  const int64_t need = sgm_unet_workspace_bytes(net, roi, batch);
  if (need < 0) return (int32_t)need;
  SGM_REQUIRE(workspace_bytes >= need, SGM_ERR_WORKSPACE, "workspace too small: need %lld bytes, got %lld",
              (long long)need, (long long)workspace_bytes);
  cudaStream_t st = (cudaStream_t)stream;
  Bump ws{(char*)workspace_dev, workspace_bytes, 0, false};
  int* org_dev = (int*)ws.take((int64_t)batch * 3 * sizeof(int));
  std::vector<int> org(batch * 3, 0);
  for (int b = 0; b < batch; ++b) org[b * 3] = b * net->cin * roi[0];
  SGM_CUDA_CHECK(cudaMemcpyAsync(org_dev, org.data(), org.size() * sizeof(int), cudaMemcpyHostToDevice, st));
  const long long vox = (long long)roi[0] * roi[1] * roi[2];
  HeadTarget head;
  memset(&head, 0, sizeof(head));
  head.kind = OUT_PLANAR, head.out = logits_dev, head.cstride = vox, head.nstride = vox * net->cout;
  net->last_launches = 0;
  return run_network(net, x_dev, vox, roi[1], roi[2], org_dev, batch, roi, ws, head, st, false);
}

static int check_cfg(const sgm_unet* net, const sgm_sw_cfg* cfg) {
  SGM_REQUIRE(cfg, SGM_ERR_INVALID, "null sgm_sw_cfg");
  for (int a = 0; a < 3; ++a) {
    SGM_REQUIRE(cfg->roi[a] >= 1 && cfg->roi[a] <= 512 && cfg->dims[a] >= cfg->roi[a], SGM_ERR_INVALID,
                "axis %d: roi %d / dims %d invalid (pad the volume to at least the roi)", a, cfg->roi[a],
                cfg->dims[a]);
    SGM_REQUIRE(cfg->n_starts[a] >= 1 && cfg->n_starts[a] <= SGM_MAX_STARTS, SGM_ERR_INVALID,
                "axis %d: n_starts %d out of range", a, cfg->n_starts[a]);
    for (int j = 0; j < cfg->n_starts[a]; ++j)
      SGM_REQUIRE(cfg->starts[a][j] >= 0 && cfg->starts[a][j] + cfg->roi[a] <= cfg->dims[a], SGM_ERR_INVALID,
                  "axis %d: window start %d outside the volume", a, cfg->starts[a][j]);
    SGM_REQUIRE(cfg->imap[a], SGM_ERR_INVALID, "null importance table");
  }
  SGM_REQUIRE(cfg->acc_nx >= 1 && cfg->acc_x0 >= 0 && cfg->acc_x0 + cfg->acc_nx <= cfg->dims[0],
              SGM_ERR_INVALID, "bad accumulator plane range");
  if (net) return check_roi(net, cfg->roi);
  return SGM_OK;
}

extern "C" int64_t sgm_sw_workspace_bytes(const sgm_unet* net, const sgm_sw_cfg* cfg) {
  if (!net || check_cfg(net, cfg)) return SGM_ERR_INVALID;
  const int64_t nwin = (int64_t)(cfg->a0_end - cfg->a0_begin) * cfg->n_starts[1] * cfg->n_starts[2];
  const int B = std::max(1, std::min<int>(cfg->sw_batch, (int)std::max<int64_t>(nwin, 1)));
  const int64_t net_bytes = sgm_unet_workspace_bytes(net, cfg->roi, B);
  if (net_bytes < 0) return net_bytes;
  return net_bytes + nwin * 3 * (int64_t)sizeof(int) + 3 * 512 * (int64_t)sizeof(float) + 1024;
}

extern "C" int32_t sgm_sw_accumulate(sgm_unet* net, const float* vol_dev, const sgm_sw_cfg* cfg,
                                     float* acc_dev, void* workspace_dev, int64_t workspace_bytes,
                                     void* stream) {
  SGM_REQUIRE(net && vol_dev && acc_dev && workspace_dev, SGM_ERR_INVALID, "sgm_sw_accumulate: null argument");
  int rc = check_cfg(net, cfg);
  if (rc) return rc;
  SGM_REQUIRE(cfg->a0_begin >= 0 && cfg->a0_end <= cfg->n_starts[0] && cfg->a0_begin <= cfg->a0_end,
              SGM_ERR_INVALID, "bad axis-0 start range");
  const int64_t need = sgm_sw_workspace_bytes(net, cfg);
  SGM_REQUIRE(need >= 0 && workspace_bytes >= need, SGM_ERR_WORKSPACE,
              "workspace too small: need %lld bytes, got %lld", (long long)need, (long long)workspace_bytes);
  cudaStream_t st = (cudaStream_t)stream;
  // window list in MONAI order (axis 0 slowest)
  std::vector<int> org_vol, org_acc;
  for (int a0 = cfg->a0_begin; a0 < cfg->a0_end; ++a0) {
    const int s0 = cfg->starts[0][a0];
    SGM_REQUIRE(s0 >= cfg->vol_x0 && s0 + cfg->roi[0] <= cfg->vol_x0 + cfg->vol_nx, SGM_ERR_INVALID,
                "window start %d needs planes outside vol_dev [%d,%d)", s0, cfg->vol_x0, cfg->vol_x0 + cfg->vol_nx);
    for (int a1 = 0; a1 < cfg->n_starts[1]; ++a1)
      for (int a2 = 0; a2 < cfg->n_starts[2]; ++a2) {
        org_vol.push_back(s0 - cfg->vol_x0), org_vol.push_back(cfg->starts[1][a1]), org_vol.push_back(cfg->starts[2][a2]);
        org_acc.push_back(s0 - cfg->acc_x0), org_acc.push_back(cfg->starts[1][a1]), org_acc.push_back(cfg->starts[2][a2]);
      }
  }
  const int nwin = (int)(org_vol.size() / 3);
  net->last_launches = 0;
  if (nwin == 0) return SGM_OK;
  Bump ws{(char*)workspace_dev, workspace_bytes, 0, false};
  int* org_dev = (int*)ws.take((int64_t)nwin * 3 * sizeof(int));
  float* imap_dev = (float*)ws.take(3 * 512 * sizeof(float));
  SGM_CUDA_CHECK(cudaMemcpyAsync(org_dev, org_vol.data(), org_vol.size() * sizeof(int), cudaMemcpyHostToDevice, st));
  for (int a = 0; a < 3; ++a)
    SGM_CUDA_CHECK(cudaMemcpyAsync(imap_dev + a * 512, cfg->imap[a], cfg->roi[a] * sizeof(float),
                                   cudaMemcpyHostToDevice, st));
  const int B = std::max(1, std::min(cfg->sw_batch, nwin));
  const long long plane = (long long)cfg->dims[1] * cfg->dims[2];
  HeadTarget head;
  memset(&head, 0, sizeof(head));
  head.kind = OUT_BLEND, head.out = acc_dev;
  head.cstride = (long long)cfg->acc_nx * plane;
  head.ad0 = cfg->acc_nx, head.ad1 = cfg->dims[1], head.ad2 = cfg->dims[2];
  for (int a = 0; a < 3; ++a) head.imap[a] = imap_dev + a * 512;
  head.floor = cfg->imap_floor;
  const int64_t ws_mark = ws.off;
  for (int w0 = 0; w0 < nwin; w0 += B) {
    const int nb = std::min(B, nwin - w0);
    ws.off = ws_mark;
    head.wo_host = org_acc.data() + (size_t)w0 * 3;
    rc = run_network(net, vol_dev, (long long)cfg->vol_nx * plane, cfg->dims[1], cfg->dims[2],
                     org_dev + (size_t)w0 * 3, nb, cfg->roi, ws, head, st, false);
    if (rc) return rc;
  }
  return SGM_OK;
}

// Window list of a schedule restricted to the axis-0 start range of the call, MONAI order.
static int build_windows(const sgm_sw_cfg* cfg, std::vector<int>& org_vol, std::vector<int>& org_acc) {
  for (int a0 = cfg->a0_begin; a0 < cfg->a0_end; ++a0) {
    const int s0 = cfg->starts[0][a0];
    SGM_REQUIRE(s0 >= cfg->vol_x0 && s0 + cfg->roi[0] <= cfg->vol_x0 + cfg->vol_nx, SGM_ERR_INVALID,
                "window start %d needs planes outside vol_dev [%d,%d)", s0, cfg->vol_x0, cfg->vol_x0 + cfg->vol_nx);
    for (int a1 = 0; a1 < cfg->n_starts[1]; ++a1)
      for (int a2 = 0; a2 < cfg->n_starts[2]; ++a2) {
        org_vol.push_back(s0 - cfg->vol_x0), org_vol.push_back(cfg->starts[1][a1]), org_vol.push_back(cfg->starts[2][a2]);
        org_acc.push_back(s0 - cfg->acc_x0), org_acc.push_back(cfg->starts[1][a1]), org_acc.push_back(cfg->starts[2][a2]);
      }
  }
  return SGM_OK;
}

extern "C" int64_t sgm_sw_predict_workspace_bytes(const sgm_unet* net, const sgm_sw_cfg* cfg) {
  const int64_t base = sgm_sw_workspace_bytes(net, cfg);
  if (base < 0) return base;
  const int64_t nwin = (int64_t)(cfg->a0_end - cfg->a0_begin) * cfg->n_starts[1] * cfg->n_starts[2];
  const int64_t roivox = (int64_t)cfg->roi[0] * cfg->roi[1] * cfg->roi[2];
  return base + nwin * net->cout * roivox * 4 + 3 * SGM_MAX_STARTS * 4 + 1024;
}

extern "C" int32_t sgm_sw_predict(sgm_unet* net, const float* vol_dev, const sgm_sw_cfg* cfg, float* logits_dev,
                                  uint8_t* labels_dev, float* probs_dev, void* workspace_dev,
                                  int64_t workspace_bytes, void* stream) {
  SGM_REQUIRE(net && vol_dev && workspace_dev, SGM_ERR_INVALID, "sgm_sw_predict: null argument");
  int rc = check_cfg(net, cfg);
  if (rc) return rc;
  SGM_REQUIRE(cfg->a0_begin >= 0 && cfg->a0_end <= cfg->n_starts[0] && cfg->a0_begin <= cfg->a0_end,
              SGM_ERR_INVALID, "bad axis-0 start range");
  const int64_t need = sgm_sw_predict_workspace_bytes(net, cfg);
  SGM_REQUIRE(need >= 0 && workspace_bytes >= need, SGM_ERR_WORKSPACE,
              "workspace too small: need %lld bytes, got %lld", (long long)need, (long long)workspace_bytes);
  cudaStream_t st = (cudaStream_t)stream;
  std::vector<int> org_vol, org_acc;
  rc = build_windows(cfg, org_vol, org_acc);
  if (rc) return rc;
  const int nwin = (int)(org_vol.size() / 3);
  net->last_launches = 0;
  SGM_REQUIRE(nwin > 0, SGM_ERR_INVALID, "sgm_sw_predict: no windows in the axis-0 range");
  Bump ws{(char*)workspace_dev, workspace_bytes, 0, false};
  int* org_dev = (int*)ws.take((int64_t)nwin * 3 * sizeof(int));
  float* imap_dev = (float*)ws.take(3 * 512 * sizeof(float));
  int* starts_dev = (int*)ws.take(3 * SGM_MAX_STARTS * sizeof(int));
  const long long roivox = (long long)cfg->roi[0] * cfg->roi[1] * cfg->roi[2];
  float* wl = (float*)ws.take((int64_t)nwin * net->cout * roivox * 4);
  SGM_CUDA_CHECK(cudaMemcpyAsync(org_dev, org_vol.data(), org_vol.size() * sizeof(int), cudaMemcpyHostToDevice, st));
  SGM_CUDA_CHECK(cudaMemcpyAsync(starts_dev, cfg->starts, 3 * SGM_MAX_STARTS * sizeof(int), cudaMemcpyHostToDevice, st));
  for (int a = 0; a < 3; ++a)
    SGM_CUDA_CHECK(cudaMemcpyAsync(imap_dev + a * 512, cfg->imap[a], cfg->roi[a] * sizeof(float),
                                   cudaMemcpyHostToDevice, st));
  const int B = std::max(1, std::min(cfg->sw_batch, nwin));
  const long long plane = (long long)cfg->dims[1] * cfg->dims[2];
  HeadTarget head;
  memset(&head, 0, sizeof(head));
  head.kind = OUT_PLANAR, head.weighted = 1;
  head.cstride = roivox, head.nstride = roivox * net->cout;
  for (int a = 0; a < 3; ++a) head.imap[a] = imap_dev + a * 512;
  head.floor = cfg->imap_floor;
  const int64_t ws_mark = ws.off;
  int64_t launches = 0;
  for (int w0 = 0; w0 < nwin; w0 += B) {
    const int nb = std::min(B, nwin - w0);
    ws.off = ws_mark;
    head.out = wl + (size_t)w0 * net->cout * roivox;
    rc = run_network(net, vol_dev, (long long)cfg->vol_nx * plane, cfg->dims[1], cfg->dims[2],
                     org_dev + (size_t)w0 * 3, nb, cfg->roi, ws, head, st, false);
    if (rc) return rc;
    launches += net->last_launches;
    net->last_launches = 0;
  }
  const float* imaps[3] = {imap_dev, imap_dev + 512, imap_dev + 1024};
  cudaEvent_t pe = nullptr;
  if (net->profiling) {  // the blend kernel is profile slot n_convs
    sgm_unet::ProfEv e;
    e.a = net->ev_get(), e.b = net->ev_get(), e.conv = (int)net->convs.size();
    cudaEventRecord(e.a, st);
    pe = e.b;
    net->prof_pending.push_back(e);
  }
  rc = launch_gather_blend(wl, net->cout, cfg, starts_dev, imaps, logits_dev, labels_dev, probs_dev, st);
  if (pe) cudaEventRecord(pe, st);
  net->last_launches = launches + 1;
  return rc;
}

extern "C" int64_t sgm_sw_windows_workspace_bytes(const sgm_unet* net, const sgm_sw_cfg* cfg) {
  if (!net || check_cfg(net, cfg)) return SGM_ERR_INVALID;
  const int B = std::max(1, cfg->sw_batch);
  const int64_t net_bytes = sgm_unet_workspace_bytes(net, cfg->roi, B);
  if (net_bytes < 0) return net_bytes;
  const int64_t nwin = (int64_t)cfg->n_starts[0] * cfg->n_starts[1] * cfg->n_starts[2];
  return net_bytes + nwin * 3 * (int64_t)sizeof(int) + 3 * 512 * (int64_t)sizeof(float) + 2048;
}

extern "C" int32_t sgm_sw_windows(sgm_unet* net, const float* vol_dev, const sgm_sw_cfg* cfg, int64_t w_first,
                                  int64_t w_count, float* wl_dev, void* workspace_dev, int64_t workspace_bytes,
                                  void* stream) {
  SGM_REQUIRE(net && vol_dev && wl_dev && workspace_dev, SGM_ERR_INVALID, "sgm_sw_windows: null argument");
  int rc = check_cfg(net, cfg);
  if (rc) return rc;
  const int n1 = cfg->n_starts[1], n2 = cfg->n_starts[2];
  const int64_t nwin_all = (int64_t)cfg->n_starts[0] * n1 * n2;
  SGM_REQUIRE(w_first >= 0 && w_count >= 0 && w_first + w_count <= nwin_all, SGM_ERR_INVALID,
              "sgm_sw_windows: window range [%lld, %lld) outside the schedule's %lld windows", (long long)w_first,
              (long long)(w_first + w_count), (long long)nwin_all);
  net->last_launches = 0;
  if (w_count == 0) return SGM_OK;
  const int64_t need = sgm_sw_windows_workspace_bytes(net, cfg);
  SGM_REQUIRE(need >= 0 && workspace_bytes >= need, SGM_ERR_WORKSPACE, "workspace too small: need %lld bytes, got %lld",
              (long long)need, (long long)workspace_bytes);
  cudaStream_t st = (cudaStream_t)stream;
  std::vector<int> org_vol;
  for (int64_t w = w_first; w < w_first + w_count; ++w) {
    const int a0 = (int)(w / (n1 * n2)), a1 = (int)((w / n2) % n1), a2 = (int)(w % n2);
    const int s0 = cfg->starts[0][a0];
    SGM_REQUIRE(s0 >= cfg->vol_x0 && s0 + cfg->roi[0] <= cfg->vol_x0 + cfg->vol_nx, SGM_ERR_INVALID,
                "window start %d needs planes outside vol_dev [%d,%d)", s0, cfg->vol_x0, cfg->vol_x0 + cfg->vol_nx);
    org_vol.push_back(s0 - cfg->vol_x0), org_vol.push_back(cfg->starts[1][a1]), org_vol.push_back(cfg->starts[2][a2]);
  }
  const int nwin = (int)w_count;
  Bump ws{(char*)workspace_dev, workspace_bytes, 0, false};
  int* org_dev = (int*)ws.take((int64_t)nwin * 3 * sizeof(int));
  float* imap_dev = (float*)ws.take(3 * 512 * sizeof(float));
  SGM_CUDA_CHECK(cudaMemcpyAsync(org_dev, org_vol.data(), org_vol.size() * sizeof(int), cudaMemcpyHostToDevice, st));
  for (int a = 0; a < 3; ++a)
    SGM_CUDA_CHECK(cudaMemcpyAsync(imap_dev + a * 512, cfg->imap[a], cfg->roi[a] * sizeof(float),
                                   cudaMemcpyHostToDevice, st));
  // (pageable sources: cudaMemcpyAsync has staged them when it returns)
  const int B = std::max(1, std::min(cfg->sw_batch, nwin));
  const long long plane = (long long)cfg->dims[1] * cfg->dims[2];
  const long long roivox = (long long)cfg->roi[0] * cfg->roi[1] * cfg->roi[2];
  HeadTarget head;
  memset(&head, 0, sizeof(head));
  head.kind = OUT_PLANAR, head.weighted = 1;
  head.cstride = roivox, head.nstride = roivox * net->cout;
  for (int a = 0; a < 3; ++a) head.imap[a] = imap_dev + a * 512;
  head.floor = cfg->imap_floor;
  const int64_t ws_mark = ws.off;
  int64_t launches = 0;
  for (int w0 = 0; w0 < nwin; w0 += B) {
    const int nb = std::min(B, nwin - w0);
    ws.off = ws_mark;
    head.out = wl_dev + (size_t)w0 * net->cout * roivox;
    rc = run_network(net, vol_dev, (long long)cfg->vol_nx * plane, cfg->dims[1], cfg->dims[2],
                     org_dev + (size_t)w0 * 3, nb, cfg->roi, ws, head, st, false);
    if (rc) return rc;
    launches += net->last_launches;
    net->last_launches = 0;
  }
  net->last_launches = launches;
  return SGM_OK;
}

extern "C" int32_t sgm_sw_blend(const sgm_sw_cfg* cfg, int32_t channels, const float* wl_dev, float* logits_dev,
                                uint8_t* labels_dev, float* probs_dev, void* scratch_dev, void* stream) {
  SGM_REQUIRE(wl_dev && scratch_dev && channels >= 1, SGM_ERR_INVALID, "sgm_sw_blend: bad argument");
  int rc = check_cfg(nullptr, cfg);
  if (rc) return rc;
  SGM_REQUIRE(cfg->a0_begin >= 0 && cfg->a0_end <= cfg->n_starts[0] && cfg->a0_begin < cfg->a0_end, SGM_ERR_INVALID,
              "bad axis-0 start range");
  cudaStream_t st = (cudaStream_t)stream;
  int* starts_dev = (int*)scratch_dev;                                   // 3 * SGM_MAX_STARTS ints
  float* imap_dev = (float*)((char*)scratch_dev + 3 * SGM_MAX_STARTS * sizeof(int) + 256);
  SGM_CUDA_CHECK(cudaMemcpyAsync(starts_dev, cfg->starts, 3 * SGM_MAX_STARTS * sizeof(int), cudaMemcpyHostToDevice, st));
  for (int a = 0; a < 3; ++a)
    SGM_CUDA_CHECK(cudaMemcpyAsync(imap_dev + a * 512, cfg->imap[a], cfg->roi[a] * sizeof(float),
                                   cudaMemcpyHostToDevice, st));
  const float* imaps[3] = {imap_dev, imap_dev + 512, imap_dev + 1024};
  return launch_gather_blend(wl_dev, channels, cfg, starts_dev, imaps, logits_dev, labels_dev, probs_dev, st);
}

extern "C" int32_t sgm_sw_finalize(const float* acc_dev, int32_t channels, const sgm_sw_cfg* cfg,
                                   float* logits_dev, uint8_t* labels_dev, float* probs_dev, void* stream) {
  SGM_REQUIRE(acc_dev && channels >= 1, SGM_ERR_INVALID, "sgm_sw_finalize: bad argument");
  int rc = check_cfg(nullptr, cfg);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  // per-call tables from the stream-ordered allocator: concurrent finalisations (other networks / streams on the same
  // device, different schedules) never share them
  const size_t bytes = 3 * SGM_MAX_STARTS * sizeof(int) + 3 * 512 * sizeof(float);
  char* tab = nullptr;
  SGM_CUDA_CHECK(cudaMallocAsync((void**)&tab, bytes, st));
  int* starts_ptr = (int*)tab;
  float* imap_ptr = (float*)(tab + 3 * SGM_MAX_STARTS * sizeof(int));
  cudaError_t e = cudaMemcpyAsync(starts_ptr, cfg->starts, sizeof(int) * 3 * SGM_MAX_STARTS, cudaMemcpyHostToDevice, st);
  for (int a = 0; a < 3 && e == cudaSuccess; ++a)
    e = cudaMemcpyAsync(imap_ptr + 512 * a, cfg->imap[a], sizeof(float) * cfg->roi[a], cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) {
    const float* imaps[3] = {imap_ptr, imap_ptr + 512, imap_ptr + 1024};
    rc = launch_finalize(acc_dev, channels, cfg, starts_ptr, imaps, logits_dev, labels_dev, probs_dev, st);
  }
  cudaFreeAsync(tab, st);
  if (e != cudaSuccess) {
    set_error("sgm_sw_finalize: table upload failed: %s", cudaGetErrorString(e));
    return SGM_ERR_CUDA;
  }
  return rc;
}

extern "C" int32_t sgm_unet_check(sgm_unet* net, void* stream) {
  SGM_REQUIRE(net, SGM_ERR_INVALID, "sgm_unet_check: null handle");
  SGM_CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
  if (!net->err_dev) return SGM_OK;
  int flag = 0;
  SGM_CUDA_CHECK(cudaMemcpy(&flag, net->err_dev, sizeof(int), cudaMemcpyDeviceToHost));
  if (flag) {
    cudaMemset(net->err_dev, 0, sizeof(int));
    set_error("tcgen05 conv pipeline timed out (wait code %d: 1 weight ring, 2 TMEM drain, 3 weights, 4 accumulator; 2x plane "
              "sweep; row sweep: 41 ring slot, 42 weights, 43 plane landed, 44 row cleared, 45 residual plane, 46 row multiplied)",
              flag);
    return SGM_ERR_CUDA;
  }
  return SGM_OK;
}

extern "C" int32_t sgm_debug_conv(sgm_unet* net, int32_t conv_index, int32_t use_tc, int32_t fused,
                                  const void* in0, int32_t cg0, const void* in1, int32_t cg1, const void* res,
                                  void* out, void* out2, int32_t n, const int32_t in_dims[3], int32_t out_dims[3],
                                  void* stream) {
  SGM_REQUIRE(net && conv_index >= 0 && conv_index < (int)net->convs.size() && in0 && out && out_dims,
              SGM_ERR_INVALID, "sgm_debug_conv: bad argument");
  const PackedConv& pc = net->convs[conv_index];
  SGM_REQUIRE(pc.kind != SGM_KIND_IDENTITY, SGM_ERR_INVALID, "sgm_debug_conv: identity layer");
  const bool bf16 = net->precision == SGM_PRECISION_BF16;
  cudaStream_t st = (cudaStream_t)stream;
  const bool tr2 = pc.kind == SGM_KIND_CONV_TRANSPOSE && pc.s[1] == 2;
  for (int i = 0; i < 3; ++i)
    out_dims[i] = tr2 ? in_dims[i] * pc.s[i] : out_dim(in_dims[i], pc.k[i], pc.s[i], pc.pad[i]);
  SGM_REQUIRE(cg0 + cg1 == pc.cgin, SGM_ERR_INVALID, "sgm_debug_conv: expected %d input channel groups", pc.cgin);
  if (use_tc) {
    const tc::TcConv* t = fused ? pc.tc_fused : pc.tc;
    SGM_REQUIRE(bf16 && t, SGM_ERR_UNSUPPORTED, "sgm_debug_conv: no tcgen05 packing for conv %d", conv_index);
    tc::TcIO io;
    io.in0 = in0, io.cg0 = cg0, io.in1 = in1, io.cg1 = cg1, io.n = n;
    for (int i = 0; i < 3; ++i) io.id[i] = in_dims[i], io.od[i] = out_dims[i];
    io.outA = out, io.cgA = pc.cgout, io.res = res;
    if (fused) {
      SGM_REQUIRE(out2, SGM_ERR_INVALID, "sgm_debug_conv: fused conv needs out2");
      io.outB = out2, io.cgB = net->convs[conv_index + 2].cgout;
    }
    return tc::tc_launch(*t, io, net->err_dev, st);
  }
  ConvArgs a;
  memset(&a, 0, sizeof(a));
  a.in0 = in0, a.cg0 = cg0, a.in1 = in1, a.cg1 = cg1, a.cin_real = pc.cin, a.n = n;
  for (int i = 0; i < 3; ++i) {
    a.id[i] = in_dims[i], a.od[i] = out_dims[i];
    a.k[i] = pc.k[i], a.s[i] = pc.s[i], a.pad[i] = pc.pad[i];
  }
  a.w = pc.w32, a.bias = pc.bias, a.cout_groups = pc.cgout, a.act = pc.act, a.alpha = pc.alpha, a.c_real = pc.cout;
  a.out = out, a.res = res;
  if (tr2) return launch_convT_fp32(a, bf16, st);
  return launch_conv_fp32(a, bf16, false, OUT_CG8, st);
}

extern "C" int32_t sgm_unet_set_profiling(sgm_unet* net, int32_t on) {
  SGM_REQUIRE(net, SGM_ERR_INVALID, "sgm_unet_set_profiling: null handle");
  net->profiling = on != 0;
  return SGM_OK;
}

extern "C" int32_t sgm_unet_get_profile(sgm_unet* net, double* ms, int64_t* launches, int32_t n, void* stream) {
  SGM_REQUIRE(net && ms && launches && n == (int)net->convs.size() + 1, SGM_ERR_INVALID,
              "sgm_unet_get_profile: expected arrays of %d entries (n_convs + 1: the last is the blend kernel)",
              net ? (int)net->convs.size() + 1 : 0);
  SGM_CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
  net->prof_ms.resize(n, 0.0);
  net->prof_n.resize(n, 0);
  for (auto& e : net->prof_pending) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, e.a, e.b) == cudaSuccess) {
      net->prof_ms[e.conv] += t;
      net->prof_n[e.conv] += 1;
    }
    net->ev_pool.push_back(e.a);
    net->ev_pool.push_back(e.b);
  }
  net->prof_pending.clear();
  for (int i = 0; i < n; ++i) {
    ms[i] = net->prof_ms[i], launches[i] = net->prof_n[i];
    net->prof_ms[i] = 0.0, net->prof_n[i] = 0;
  }
  return SGM_OK;
}
