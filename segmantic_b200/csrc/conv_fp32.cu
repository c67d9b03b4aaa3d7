// CUDA-core convolution family (FFMA, fp32 accumulate), templated on the CG8 storage type:
//   float          the `precision="fp32"` path (reference numerics: the reference runs fp32 always,
//                  seg/monai_unet.py:551-670);
//   __nv_bfloat16  the on-device cross-check of the tcgen05 family (same storage, same rounding
//                  points) and the executor for layer shapes the tcgen05 family does not cover.
// Direct (gather-form) convolution over CG8 activations.
//
//   conv_fp32_kernel   k in {1,3}^3, stride in {1,2}^3, zero padding k/2 at the WINDOW border
//                      (each ROI window is convolved in isolation, as MONAI does).
//                      IN_PLANAR: the stem reads window n straight from the planar volume.
//                      OUT_BLEND / OUT_PLANAR: the head writes importance-weighted logits into the
//                      volume accumulator / plain planar logits.
//   convT_fp32_kernel  ConvTranspose k3 s2 p1 op1 as 8 output-parity classes per input voxel
//                      (1,2,2,2,4,4,4,8 taps), input = channel concat of two tensors (the skip
//                      connection's torch.cat is never materialised).
#include "common.cuh"

namespace sgm {

namespace {

constexpr int CO_T = 16;   // couts per thread, conv
constexpr int CO_TT = 8;   // couts per thread, transposed conv
constexpr int PT = 2;      // output positions per thread (along d0), conv

__device__ __forceinline__ float prelu(float v, float alpha) { return v > 0.f ? v : alpha * v; }

// One CG8 voxel-group (8 channels): 32 B in fp32 storage, 16 B in bf16 storage.
__device__ __forceinline__ void load8(const float* p, float x[8]) {
  const float4 lo = __ldg(reinterpret_cast<const float4*>(p));
  const float4 hi = __ldg(reinterpret_cast<const float4*>(p) + 1);
  x[0] = lo.x, x[1] = lo.y, x[2] = lo.z, x[3] = lo.w, x[4] = hi.x, x[5] = hi.y, x[6] = hi.z, x[7] = hi.w;
}
__device__ __forceinline__ void load8(const __nv_bfloat16* p, float x[8]) {
  const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    x[2 * i] = __uint_as_float(w[i] << 16);
    x[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ void store8(float* p, const float v[8]) {
  reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float v[8]) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat162 b = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    w[i] = *reinterpret_cast<const uint32_t*>(&b);
  }
  *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
}

template <typename T, bool IN_PLANAR, int OUT_KIND>
__global__ void __launch_bounds__(128) conv_fp32_kernel(const ConvArgs a) {
  __shared__ __align__(16) float wsm[27 * 8 * CO_T];
  const int tid = threadIdx.x;
  const int t2 = tid & 7, t1 = (tid >> 3) & 3, t0 = tid >> 5;
  const int nt2 = (a.od[2] + 7) >> 3, nt1 = (a.od[1] + 3) >> 2;
  int tile = blockIdx.x;
  const int b2 = tile % nt2;
  tile /= nt2;
  const int b1 = tile % nt1;
  const int b0 = tile / nt1;
  const int o2 = b2 * 8 + t2, o1 = b1 * 4 + t1;
  int o0[PT];
  bool ovalid[PT];
#pragma unroll
  for (int p = 0; p < PT; ++p) {
    o0[p] = b0 * (4 * PT) + t0 + 4 * p;
    ovalid[p] = o0[p] < a.od[0] && o1 < a.od[1] && o2 < a.od[2];
  }
  const int n = blockIdx.z, coblk = blockIdx.y;
  const int ntaps = a.k[0] * a.k[1] * a.k[2];
  const int cgin = a.cg0 + a.cg1;
  const long long ivox = (long long)a.id[0] * a.id[1] * a.id[2];

  float acc[PT][CO_T];
#pragma unroll
  for (int p = 0; p < PT; ++p)
#pragma unroll
    for (int c = 0; c < CO_T; ++c) acc[p][c] = 0.f;

  int worg[3] = {0, 0, 0};
  if (IN_PLANAR) {
    worg[0] = a.win_origin[n * 3 + 0];
    worg[1] = a.win_origin[n * 3 + 1];
    worg[2] = a.win_origin[n * 3 + 2];
  }

  for (int cg = 0; cg < cgin; ++cg) {
    const float4* wsrc =
        reinterpret_cast<const float4*>(reinterpret_cast<const float*>(a.w) +
                                        ((size_t)(coblk * cgin + cg) * ntaps * 8 * CO_T));
    for (int i = tid; i < ntaps * 8 * CO_T / 4; i += 128) reinterpret_cast<float4*>(wsm)[i] = wsrc[i];
    __syncthreads();
    const T* src = nullptr;
    if (!IN_PLANAR) {
      src = (cg < a.cg0)
                ? reinterpret_cast<const T*>(a.in0) + ((long long)n * a.cg0 + cg) * ivox * 8
                : reinterpret_cast<const T*>(a.in1) + ((long long)n * a.cg1 + (cg - a.cg0)) * ivox * 8;
    }
    int tap = 0;
    for (int k0 = 0; k0 < a.k[0]; ++k0)
      for (int k1 = 0; k1 < a.k[1]; ++k1)
        for (int k2 = 0; k2 < a.k[2]; ++k2, ++tap) {
          const int i1 = o1 * a.s[1] + k1 - a.pad[1];
          const int i2 = o2 * a.s[2] + k2 - a.pad[2];
          const bool v12 = i1 >= 0 && i1 < a.id[1] && i2 >= 0 && i2 < a.id[2];
          float x[PT][8];
#pragma unroll
          for (int p = 0; p < PT; ++p) {
            const int i0 = o0[p] * a.s[0] + k0 - a.pad[0];
            const bool v = ovalid[p] && v12 && i0 >= 0 && i0 < a.id[0];
            if (IN_PLANAR) {
#pragma unroll
              for (int ci = 0; ci < 8; ++ci) x[p][ci] = 0.f;
              if (v) {
                const long long off =
                    ((long long)(worg[0] + i0) * a.vd1 + (worg[1] + i1)) * a.vd2 + (worg[2] + i2);
                const float* vol = reinterpret_cast<const float*>(a.in0);
                for (int ci = 0; ci < 8; ++ci) {
                  const int c = cg * 8 + ci;
                  if (c < a.cin_real) x[p][ci] = __ldg(vol + c * a.vol_cstride + off);
                }
              }
            } else {
#pragma unroll
              for (int ci = 0; ci < 8; ++ci) x[p][ci] = 0.f;
              if (v) load8(src + (((long long)i0 * a.id[1] + i1) * a.id[2] + i2) * 8, x[p]);
            }
          }
          const float* wt = wsm + tap * 8 * CO_T;
#pragma unroll
          for (int ci = 0; ci < 8; ++ci) {
#pragma unroll
            for (int c4 = 0; c4 < CO_T / 4; ++c4) {
              const float4 w = *reinterpret_cast<const float4*>(wt + ci * CO_T + c4 * 4);
#pragma unroll
              for (int p = 0; p < PT; ++p) {
                acc[p][c4 * 4 + 0] = fmaf(x[p][ci], w.x, acc[p][c4 * 4 + 0]);
                acc[p][c4 * 4 + 1] = fmaf(x[p][ci], w.y, acc[p][c4 * 4 + 1]);
                acc[p][c4 * 4 + 2] = fmaf(x[p][ci], w.z, acc[p][c4 * 4 + 2]);
                acc[p][c4 * 4 + 3] = fmaf(x[p][ci], w.w, acc[p][c4 * 4 + 3]);
              }
            }
          }
        }
    __syncthreads();
  }

  // ---- epilogue: + bias, PReLU, + residual, store
  const long long ovox = (long long)a.od[0] * a.od[1] * a.od[2];
#pragma unroll
  for (int p = 0; p < PT; ++p) {
    if (!ovalid[p]) continue;
    const long long opos = ((long long)o0[p] * a.od[1] + o1) * a.od[2] + o2;
    float imw = 1.f;
    long long pl_off = 0;
    bool pl_ok = true;
    if (OUT_KIND == OUT_BLEND) {
      const int g0 = a.wo[0] + o0[p];
      pl_ok = g0 >= 0 && g0 < a.ad0;
      pl_off = ((long long)g0 * a.ad1 + (a.wo[1] + o1)) * a.ad2 + (a.wo[2] + o2);
      imw = fmaxf(__fmul_rn(__fmul_rn(a.imap0[o0[p]], a.imap1[o1]), a.imap2[o2]), a.imap_floor);
    } else if (OUT_KIND == OUT_PLANAR) {
      pl_off = (long long)n * a.pl_nstride + opos;
      if (a.pl_weighted)
        imw = fmaxf(__fmul_rn(__fmul_rn(a.imap0[o0[p]], a.imap1[o1]), a.imap2[o2]), a.imap_floor);
    }
#pragma unroll
    for (int g = 0; g < CO_T / 8; ++g) {
      const int cgo = coblk * (CO_T / 8) + g;
      if (cgo >= a.cout_groups) continue;
      float v[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float t = acc[p][g * 8 + c] + __ldg(a.bias + cgo * 8 + c);
        if (a.act) t = prelu(t, a.alpha);
        v[c] = t;
      }
      if (a.res) {
        float r[8];
        load8(reinterpret_cast<const T*>(a.res) + (((long long)n * a.cout_groups + cgo) * ovox + opos) * 8, r);
#pragma unroll
        for (int c = 0; c < 8; ++c) v[c] += r[c];
      }
      if (OUT_KIND == OUT_CG8) {
        store8(reinterpret_cast<T*>(a.out) + (((long long)n * a.cout_groups + cgo) * ovox + opos) * 8, v);
      } else if (pl_ok) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const int ch = cgo * 8 + c;
          if (ch < a.c_real) {
            float* dst = a.pl_out + ch * a.pl_cstride + pl_off;
            if (OUT_KIND == OUT_BLEND)
              *dst = __fadd_rn(*dst, __fmul_rn(v[c], imw));  // seg *= w; out += seg (two roundings)
            else
              *dst = a.pl_weighted ? __fmul_rn(v[c], imw) : v[c];
          }
        }
      }
    }
  }
}

// Per-axis tap of a stride-2 k3 p1 transposed conv: output parity class p, input shift sh -> kernel
// index, or -1.  o = 2j + p reads in[j + sh] * W[k] with o + 1 - k = 2(j + sh).
__device__ __forceinline__ int tconv_tap(int p, int sh) {
  return p == 0 ? (sh == 0 ? 1 : -1) : (sh == 0 ? 2 : 0);
}

template <typename T, bool FLAT0>
__global__ void __launch_bounds__(128) convT_fp32_kernel(const ConvArgs a) {
  __shared__ __align__(16) float wsm[27 * 8 * CO_TT];
  constexpr int NC0 = FLAT0 ? 1 : 2;
  constexpr int NCLS = NC0 * 4;
  const int tid = threadIdx.x;
  const int t2 = tid & 7, t1 = (tid >> 3) & 3, t0 = tid >> 5;
  const int nt2 = (a.id[2] + 7) >> 3, nt1 = (a.id[1] + 3) >> 2;
  int tile = blockIdx.x;
  const int b2 = tile % nt2;
  tile /= nt2;
  const int b1 = tile % nt1;
  const int b0 = tile / nt1;
  const int j2 = b2 * 8 + t2, j1 = b1 * 4 + t1, j0 = b0 * 4 + t0;
  const bool jvalid = j0 < a.id[0] && j1 < a.id[1] && j2 < a.id[2];
  const int n = blockIdx.z, coblk = blockIdx.y;
  const int ntaps = a.k[0] * a.k[1] * a.k[2];
  const int cgin = a.cg0 + a.cg1;
  const long long ivox = (long long)a.id[0] * a.id[1] * a.id[2];

  float acc[NCLS][CO_TT];
#pragma unroll
  for (int q = 0; q < NCLS; ++q)
#pragma unroll
    for (int c = 0; c < CO_TT; ++c) acc[q][c] = 0.f;

  for (int cg = 0; cg < cgin; ++cg) {
    const float4* wsrc =
        reinterpret_cast<const float4*>(reinterpret_cast<const float*>(a.w) +
                                        ((size_t)(coblk * cgin + cg) * ntaps * 8 * CO_TT));
    for (int i = tid; i < ntaps * 8 * CO_TT / 4; i += 128) reinterpret_cast<float4*>(wsm)[i] = wsrc[i];
    __syncthreads();
    const T* src =
        (cg < a.cg0) ? reinterpret_cast<const T*>(a.in0) + ((long long)n * a.cg0 + cg) * ivox * 8
                     : reinterpret_cast<const T*>(a.in1) + ((long long)n * a.cg1 + (cg - a.cg0)) * ivox * 8;
#pragma unroll
    for (int sh0 = 0; sh0 < NC0; ++sh0)
#pragma unroll
      for (int sh1 = 0; sh1 < 2; ++sh1)
#pragma unroll
        for (int sh2 = 0; sh2 < 2; ++sh2) {
          const int i0 = j0 + sh0, i1 = j1 + sh1, i2 = j2 + sh2;
          const bool v = jvalid && i0 < a.id[0] && i1 < a.id[1] && i2 < a.id[2];
          float x[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          if (v) load8(src + (((long long)i0 * a.id[1] + i1) * a.id[2] + i2) * 8, x);
#pragma unroll
          for (int p0 = 0; p0 < NC0; ++p0)
#pragma unroll
            for (int p1 = 0; p1 < 2; ++p1)
#pragma unroll
              for (int p2 = 0; p2 < 2; ++p2) {
                const int kk0 = FLAT0 ? 0 : tconv_tap(p0, sh0);
                const int kk1 = tconv_tap(p1, sh1), kk2 = tconv_tap(p2, sh2);
                if (kk0 < 0 || kk1 < 0 || kk2 < 0) continue;
                const int tap = (kk0 * 3 + kk1) * 3 + kk2;
                const int q = (p0 * 2 + p1) * 2 + p2;
                const float* wt = wsm + tap * 8 * CO_TT;
#pragma unroll
                for (int ci = 0; ci < 8; ++ci) {
                  const float4 w0 = *reinterpret_cast<const float4*>(wt + ci * CO_TT);
                  const float4 w1 = *reinterpret_cast<const float4*>(wt + ci * CO_TT + 4);
                  acc[q][0] = fmaf(x[ci], w0.x, acc[q][0]);
                  acc[q][1] = fmaf(x[ci], w0.y, acc[q][1]);
                  acc[q][2] = fmaf(x[ci], w0.z, acc[q][2]);
                  acc[q][3] = fmaf(x[ci], w0.w, acc[q][3]);
                  acc[q][4] = fmaf(x[ci], w1.x, acc[q][4]);
                  acc[q][5] = fmaf(x[ci], w1.y, acc[q][5]);
                  acc[q][6] = fmaf(x[ci], w1.z, acc[q][6]);
                  acc[q][7] = fmaf(x[ci], w1.w, acc[q][7]);
                }
              }
        }
    __syncthreads();
  }
  if (!jvalid || coblk >= a.cout_groups) return;
  const long long ovox = (long long)a.od[0] * a.od[1] * a.od[2];
  T* obase = reinterpret_cast<T*>(a.out) + ((long long)n * a.cout_groups + coblk) * ovox * 8;
#pragma unroll
  for (int p0 = 0; p0 < NC0; ++p0)
#pragma unroll
    for (int p1 = 0; p1 < 2; ++p1)
#pragma unroll
      for (int p2 = 0; p2 < 2; ++p2) {
        const int q = (p0 * 2 + p1) * 2 + p2;
        const int oo0 = FLAT0 ? j0 : 2 * j0 + p0, oo1 = 2 * j1 + p1, oo2 = 2 * j2 + p2;
        float v[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float t = acc[q][c] + __ldg(a.bias + coblk * 8 + c);
          if (a.act) t = prelu(t, a.alpha);
          v[c] = t;
        }
        store8(obase + (((long long)oo0 * a.od[1] + oo1) * a.od[2] + oo2) * 8, v);
      }
}


// ---------------------------------------------------------------------------------------------- stem
// First down block: unit0 (k3, stride s, +PReLU) and the residual-branch conv (k3 stride s, or k1) read
// the same 1..8-channel window of the planar fp32 volume; both are computed by one kernel (the window
// is fetched once, straight from the volume at the window origin -- no gather copy).  K = 27*Cin is far
// too small for the tensor cores; this is a CUDA-core FFMA kernel: one output voxel per thread (lanes
// along d2), STEM_CO fused output channels per thread, weights broadcast from shared memory.
constexpr int STEM_CO = 32;
constexpr int STEM_PT = 2;

struct StemArgs {
  const float* vol;
  long long vol_cstride;
  int vd1, vd2;
  const int* win_origin;
  int cin, n;
  int id[3], od[3], k[3], s[3], pad[3];
  const float* w;     // [coblk][tap][ci][STEM_CO]
  const float* bias;  // [coblk*STEM_CO]
  int cgA, cgB;       // channel groups of the two outputs; fused channel f < cgA*8 -> A else B
  void* outA;
  void* outB;
  int actA;
  float alphaA;
};

template <typename T>
__global__ void __launch_bounds__(128) stem_kernel(const StemArgs a) {
  extern __shared__ __align__(16) float wsm[];
  const int ntaps = a.k[0] * a.k[1] * a.k[2];
  const int coblk = blockIdx.z, n = blockIdx.y;
  const int wcount = ntaps * a.cin * STEM_CO;
  for (int i = threadIdx.x; i < wcount; i += 128) wsm[i] = __ldg(a.w + (size_t)coblk * wcount + i);
  __syncthreads();
  const int nt2 = (a.od[2] + 31) >> 5, nt1 = (a.od[1] + 3) >> 2;
  int tile = blockIdx.x;
  const int b2 = tile % nt2;
  tile /= nt2;
  const int b1 = tile % nt1;
  const int o0b = (tile / nt1) * STEM_PT;  // STEM_PT output planes per thread: weights are read once for both
  const int o2 = b2 * 32 + (threadIdx.x & 31), o1 = b1 * 4 + (threadIdx.x >> 5);
  if (o1 >= a.od[1] || o2 >= a.od[2]) return;
  const int w0 = a.win_origin[n * 3 + 0], w1 = a.win_origin[n * 3 + 1], w2 = a.win_origin[n * 3 + 2];
  float acc[STEM_PT][STEM_CO];
#pragma unroll
  for (int p = 0; p < STEM_PT; ++p)
#pragma unroll
    for (int c = 0; c < STEM_CO; ++c) acc[p][c] = 0.f;
  int tap = 0;
  for (int k0 = 0; k0 < a.k[0]; ++k0) {
    for (int k1 = 0; k1 < a.k[1]; ++k1) {
      const int i1 = o1 * a.s[1] + k1 - a.pad[1];
      for (int k2 = 0; k2 < a.k[2]; ++k2, ++tap) {
        const int i2 = o2 * a.s[2] + k2 - a.pad[2];
        const bool ok12 = i1 >= 0 && i1 < a.id[1] && i2 >= 0 && i2 < a.id[2];
        for (int ci = 0; ci < a.cin; ++ci) {
          float x[STEM_PT];
#pragma unroll
          for (int p = 0; p < STEM_PT; ++p) {
            const int i0 = (o0b + p) * a.s[0] + k0 - a.pad[0];
            const bool ok = ok12 && i0 >= 0 && i0 < a.id[0];
            const long long off = ((long long)(w0 + i0) * a.vd1 + (w1 + i1)) * a.vd2 + (w2 + i2);
            x[p] = ok ? __ldg(a.vol + ci * a.vol_cstride + off) : 0.f;
          }
          const float4* wr = reinterpret_cast<const float4*>(wsm + (tap * a.cin + ci) * STEM_CO);
#pragma unroll
          for (int c4 = 0; c4 < STEM_CO / 4; ++c4) {
            const float4 w = wr[c4];
#pragma unroll
            for (int p = 0; p < STEM_PT; ++p) {
              acc[p][c4 * 4 + 0] = fmaf(x[p], w.x, acc[p][c4 * 4 + 0]);
              acc[p][c4 * 4 + 1] = fmaf(x[p], w.y, acc[p][c4 * 4 + 1]);
              acc[p][c4 * 4 + 2] = fmaf(x[p], w.z, acc[p][c4 * 4 + 2]);
              acc[p][c4 * 4 + 3] = fmaf(x[p], w.w, acc[p][c4 * 4 + 3]);
            }
          }
        }
      }
    }
  }
  const long long ovox = (long long)a.od[0] * a.od[1] * a.od[2];
#pragma unroll
  for (int p = 0; p < STEM_PT; ++p) {
    const int o0 = o0b + p;
    if (o0 >= a.od[0]) continue;
    const long long opos = ((long long)o0 * a.od[1] + o1) * a.od[2] + o2;
#pragma unroll
    for (int g = 0; g < STEM_CO / 8; ++g) {
      const int fcg = coblk * (STEM_CO / 8) + g;  // fused channel group
      if (fcg >= a.cgA + a.cgB) continue;
      const bool isA = fcg < a.cgA;
      float v[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float t = acc[p][g * 8 + c] + __ldg(a.bias + fcg * 8 + c);
        if (isA && a.actA) t = prelu(t, a.alphaA);
        v[c] = t;
      }
      if (isA)
        store8(reinterpret_cast<T*>(a.outA) + (((long long)n * a.cgA + fcg) * ovox + opos) * 8, v);
      else
        store8(reinterpret_cast<T*>(a.outB) + (((long long)n * a.cgB + (fcg - a.cgA)) * ovox + opos) * 8, v);
    }
  }
}

}  // namespace

int stem_cout_tile() { return STEM_CO; }

int launch_stem(const ConvArgs& a, int cgA, int cgB, void* outA, void* outB, bool bf16_storage, cudaStream_t st) {
  StemArgs s;
  s.vol = reinterpret_cast<const float*>(a.in0), s.vol_cstride = a.vol_cstride, s.vd1 = a.vd1, s.vd2 = a.vd2;
  s.win_origin = a.win_origin, s.cin = a.cin_real, s.n = a.n;
  for (int i = 0; i < 3; ++i) s.id[i] = a.id[i], s.od[i] = a.od[i], s.k[i] = a.k[i], s.s[i] = a.s[i], s.pad[i] = a.pad[i];
  s.w = reinterpret_cast<const float*>(a.w), s.bias = a.bias;
  s.cgA = cgA, s.cgB = cgB, s.outA = outA, s.outB = outB, s.actA = a.act, s.alphaA = a.alpha;
  const int ntaps = a.k[0] * a.k[1] * a.k[2];
  const size_t smem = (size_t)ntaps * a.cin_real * STEM_CO * sizeof(float);
  SGM_REQUIRE(smem <= 48 * 1024, SGM_ERR_UNSUPPORTED, "stem: %d input channels need %zu B of shared memory", a.cin_real, smem);
  dim3 grid(ceil_div(a.od[2], 32) * ceil_div(a.od[1], 4) * ceil_div(a.od[0], STEM_PT), a.n, ceil_div((cgA + cgB) * 8, STEM_CO));
  if (bf16_storage)
    stem_kernel<__nv_bfloat16><<<grid, 128, smem, st>>>(s);
  else
    stem_kernel<float><<<grid, 128, smem, st>>>(s);
  SGM_CUDA_CHECK(cudaGetLastError());
  return SGM_OK;
}

namespace {
}  // namespace

int fp32_conv_cout_tile(bool transposed) { return transposed ? CO_TT : CO_T; }

template <typename T>
static int launch_conv_t(const ConvArgs& a, bool in_planar, int out_kind, cudaStream_t st) {
  const int nt = ceil_div(a.od[2], 8) * ceil_div(a.od[1], 4) * ceil_div(a.od[0], 4 * PT);
  dim3 grid(nt, ceil_div(a.cout_groups * 8, CO_T), a.n);
  if (in_planar) {
    SGM_REQUIRE(out_kind == OUT_CG8, SGM_ERR_UNSUPPORTED, "planar-in conv writes CG8 only");
    conv_fp32_kernel<T, true, OUT_CG8><<<grid, 128, 0, st>>>(a);
  } else if (out_kind == OUT_CG8) {
    conv_fp32_kernel<T, false, OUT_CG8><<<grid, 128, 0, st>>>(a);
  } else if (out_kind == OUT_BLEND) {
    conv_fp32_kernel<T, false, OUT_BLEND><<<grid, 128, 0, st>>>(a);
  } else {
    conv_fp32_kernel<T, false, OUT_PLANAR><<<grid, 128, 0, st>>>(a);
  }
  SGM_CUDA_CHECK(cudaGetLastError());
  return SGM_OK;
}

int launch_conv_fp32(const ConvArgs& a, bool bf16_storage, bool in_planar, int out_kind, cudaStream_t st) {
  return bf16_storage ? launch_conv_t<__nv_bfloat16>(a, in_planar, out_kind, st)
                      : launch_conv_t<float>(a, in_planar, out_kind, st);
}

int launch_convT_fp32(const ConvArgs& a, bool bf16_storage, cudaStream_t st) {
  const bool flat0 = a.k[0] == 1;
  const int nt = ceil_div(a.id[2], 8) * ceil_div(a.id[1], 4) * ceil_div(a.id[0], 4);
  dim3 grid(nt, a.cout_groups, a.n);
  if (bf16_storage) {
    if (flat0)
      convT_fp32_kernel<__nv_bfloat16, true><<<grid, 128, 0, st>>>(a);
    else
      convT_fp32_kernel<__nv_bfloat16, false><<<grid, 128, 0, st>>>(a);
  } else {
    if (flat0)
      convT_fp32_kernel<float, true><<<grid, 128, 0, st>>>(a);
    else
      convT_fp32_kernel<float, false><<<grid, 128, 0, st>>>(a);
  }
  SGM_CUDA_CHECK(cudaGetLastError());
  return SGM_OK;
}

}  // namespace sgm
