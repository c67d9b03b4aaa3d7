// Final pass of the sliding-window inferer: out = acc / count, optional argmax and softmax.
//
// MONAI (sliding_window_inference, called at seg/monai_unet.py:665) keeps a second full-volume fp32
// "count" buffer that every window adds its importance map to.  The windows are a cartesian product
// of per-axis starts and each contribution is max((g0[i]*g1[j])*g2[k], floor), so the count of a
// voxel is recomputed here on the fly, summing the covering windows in MONAI's window order (axis 0
// slowest) with the same fp32 operations -- bit-identical to the accumulated buffer, with zero HBM
// traffic for it.  HBM-bound: reads 4*C bytes/voxel, writes 1 (labels) [+4*C logits] [+4*C probs].
#include "common.cuh"

#include <algorithm>
#include <stdlib.h>

namespace sgm {

namespace {

struct FinalizeArgs {
  const float* acc;
  float* logits;
  uint8_t* labels;
  float* probs;
  int channels;
  int nx, d1, d2;  // local extent (planes acc_x0 .. acc_x0+nx)
  int x0;          // global plane index of local plane 0
  int roi[3];
  int n_starts[3];
  const int* starts;  // device int[3][SGM_MAX_STARTS]
  const float* imap0;
  const float* imap1;
  const float* imap2;
  float floor;
};

// windows j in [lo, hi] cover v  (starts ascending: start[j] <= v < start[j] + roi); any number of them (overlap is
// user-settable in [0, 1): at overlap 0.9 a voxel is covered by ten windows per axis)
__device__ __forceinline__ void cover_range_g(const int* starts, int ns, int roi, int v, int& lo, int& hi) {
  lo = ns, hi = -1;
  for (int j = 0; j < ns; ++j) {
    const int s = starts[j];
    if (s <= v && v < s + roi) {
      lo = min(lo, j);
      hi = j;
    }
  }
}

template <int CMAX>
__global__ void __launch_bounds__(256) finalize_kernel(const FinalizeArgs a) {
  const long long vox = (long long)a.nx * a.d1 * a.d2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const int* st0 = a.starts;
  const int* st1 = a.starts + SGM_MAX_STARTS;
  const int* st2 = a.starts + 2 * SGM_MAX_STARTS;
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < vox; v += stride) {
    const int z = (int)(v % a.d2);
    const long long t = v / a.d2;
    const int y = (int)(t % a.d1);
    const int x = (int)(t / a.d1) + a.x0;
    int p_lo, p_hi, q_lo, q_hi, r_lo, r_hi;
    cover_range_g(st0, a.n_starts[0], a.roi[0], x, p_lo, p_hi);
    cover_range_g(st1, a.n_starts[1], a.roi[1], y, q_lo, q_hi);
    cover_range_g(st2, a.n_starts[2], a.roi[2], z, r_lo, r_hi);
    float count = 0.f;
    for (int p = p_lo; p <= p_hi; ++p) {
      const float g0 = a.imap0[x - st0[p]];
      for (int q = q_lo; q <= q_hi; ++q) {
        const float g01 = __fmul_rn(g0, a.imap1[y - st1[q]]);
        for (int r = r_lo; r <= r_hi; ++r)
          count = __fadd_rn(count, fmaxf(__fmul_rn(g01, a.imap2[z - st2[r]]), a.floor));
      }
    }
    float val[CMAX];
    float best = 0.f;
    int arg = 0;
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
      if (c < a.channels) {
        val[c] = __fdiv_rn(__ldcs(a.acc + c * vox + v), count);
        if (c == 0 || val[c] > best) {
          best = val[c];
          arg = c;
        }
      }
    }
    if (a.labels) a.labels[v] = (uint8_t)arg;
    if (a.logits) {
#pragma unroll
      for (int c = 0; c < CMAX; ++c)
        if (c < a.channels) __stcs(a.logits + c * vox + v, val[c]);
    }
    if (a.probs) {
      float sum = 0.f;
#pragma unroll
      for (int c = 0; c < CMAX; ++c)
        if (c < a.channels) {
          val[c] = expf(val[c] - best);
          sum += val[c];
        }
      const float inv = 1.f / sum;
#pragma unroll
      for (int c = 0; c < CMAX; ++c)
        if (c < a.channels) __stcs(a.probs + c * vox + v, val[c] * inv);
    }
  }
}

// Deferred ("gather") blend: every window's importance-weighted logits were stored once
// ([window][C][roi voxels] fp32); each output voxel now sums its covering windows in MONAI's window
// order (axis 0 slowest) with the same fp32 additions as the sequential `out[slice] += seg` -- bit
// identical to the read-modify-write form, but every byte is touched once by a pure streaming kernel
// (no latency-bound RMW in the conv epilogue).  Fuses count, normalise, argmax / softmax.
struct GatherArgs {
  const float* wl;
  long long win_stride, cstride;
  float* logits;
  uint8_t* labels;
  float* probs;
  int channels;
  int nx, d1, d2, x0;
  int roi[3];
  int n_starts[3];
  int a0_begin, a0_end;
  const int* starts;
  const float* imap0;
  const float* imap1;
  const float* imap2;
  float floor;
};

// One warp per channel: lane l owns VEC consecutive voxels of a z-chunk, warp w sums channel w (w + W, ...) of
// those voxels over the covering windows -- every (window, channel) read of a warp is one contiguous 128 / 512
// byte run, a thread keeps only VEC accumulators and its window loads are independent.  The windows that cover
// a coordinate are a contiguous range of the (sorted) per-axis starts, so the covering sets are two integers per
// axis: no per-thread index arrays (the first version kept them in local memory and executed 1200 instructions
// per warp and chunk).  A block owns one (x, y) row: the axis-0 / axis-1 ranges and the count map of the row are
// computed once and shared; the normalised values meet in shared memory for the argmax / softmax across channels.
// VEC = 4 needs dims[2], roi[2] and every axis-2 window start to be multiples of 4 (aligned float4).
constexpr int kGatherWarps = 16;

// windows j in [lo, hi] cover v  (starts ascending: start[j] <= v < start[j] + roi)
__device__ __forceinline__ void cover_range(const int* starts, int ns, int roi, int v, int& lo, int& hi) {
  lo = ns, hi = -1;
  for (int j = 0; j < ns; ++j) {
    const int s = starts[j];
    if (s <= v && v < s + roi) {
      lo = min(lo, j);
      hi = j;
    }
  }
}

template <int VEC>
__global__ void __launch_bounds__(kGatherWarps * 32) gather_blend_cw_kernel(const GatherArgs a, int nwarps) {
  extern __shared__ float dyn[];  // [2][channels][32 * VEC] values | [d2] count map of the row
  __shared__ int s_st[3][SGM_MAX_STARTS];
  __shared__ float s_im[3][512];
  constexpr int CH = 32 * VEC;  // voxels per z-chunk
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int C = a.channels;
  float* vals = dyn;
  float* s_count = dyn + (size_t)2 * C * CH;
  for (int i = threadIdx.x; i < 3 * SGM_MAX_STARTS; i += blockDim.x) s_st[i / SGM_MAX_STARTS][i % SGM_MAX_STARTS] = a.starts[i];
  for (int i = threadIdx.x; i < a.roi[0]; i += blockDim.x) s_im[0][i] = a.imap0[i];
  for (int i = threadIdx.x; i < a.roi[1]; i += blockDim.x) s_im[1][i] = a.imap1[i];
  for (int i = threadIdx.x; i < a.roi[2]; i += blockDim.x) s_im[2][i] = a.imap2[i];
  __syncthreads();
  const long long vox = (long long)a.nx * a.d1 * a.d2;
  const int nchunks = (a.d2 + CH - 1) / CH;
  const long long nrows = (long long)a.nx * a.d1;
  int buf = 0;
  for (long long row = blockIdx.x; row < nrows; row += gridDim.x) {
    const int y = (int)(row % a.d1);
    const int xl = (int)(row / a.d1);
    const int x = xl + a.x0;
    int p_lo, p_hi, q_lo, q_hi;
    cover_range(s_st[0], a.n_starts[0], a.roi[0], x, p_lo, p_hi);
    cover_range(s_st[1], a.n_starts[1], a.roi[1], y, q_lo, q_hi);
    // ---- count map of the row, MONAI's window order and roundings (all threads, VEC voxels each)
    for (int z = threadIdx.x * VEC; z < a.d2; z += blockDim.x * VEC) {
      int r_lo, r_hi;
      cover_range(s_st[2], a.n_starts[2], a.roi[2], z, r_lo, r_hi);
      float count[VEC];
#pragma unroll
      for (int i = 0; i < VEC; ++i) count[i] = 0.f;
      for (int p = p_lo; p <= p_hi; ++p) {
        const float g0 = s_im[0][x - s_st[0][p]];
        for (int q = q_lo; q <= q_hi; ++q) {
          const float g01 = __fmul_rn(g0, s_im[1][y - s_st[1][q]]);
          for (int r = r_lo; r <= r_hi; ++r) {
            const int l2 = z - s_st[2][r];
#pragma unroll
            for (int i = 0; i < VEC; ++i) count[i] = __fadd_rn(count[i], fmaxf(__fmul_rn(g01, s_im[2][l2 + i]), a.floor));
          }
        }
      }
#pragma unroll
      for (int i = 0; i < VEC; ++i) s_count[z + i] = count[i];
    }
    __syncthreads();
    // windows of this rank only (multi-GPU slabs blend a sub-range of the axis-0 starts)
    const int pa = max(p_lo, a.a0_begin), pb = min(p_hi, a.a0_end - 1);
    for (int ch = 0; ch < nchunks; ++ch, buf ^= 1) {
      const int z = ch * CH + lane * VEC;
      const bool zin = z < a.d2;
      int r_lo = 0, r_hi = -1;
      if (zin) cover_range(s_st[2], a.n_starts[2], a.roi[2], z, r_lo, r_hi);
      float* vb = vals + (size_t)buf * C * CH;
      for (int c = warp; c < C; c += nwarps) {
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        const float* wl_c = a.wl + (long long)c * a.cstride;
        for (int p = pa; p <= pb; ++p) {
          const long long w0 = (long long)(p - a.a0_begin) * a.n_starts[1];
          const long long off0 = (long long)(x - s_st[0][p]) * a.roi[1];
          for (int q = q_lo; q <= q_hi; ++q) {
            const float* src01 = wl_c + (w0 + q) * a.n_starts[2] * a.win_stride + (off0 + (y - s_st[1][q])) * a.roi[2] + z;
            // up to three axis-2 windows (overlap <= 0.5) are loaded before the first add; the adds keep MONAI's order
            float4 v[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
              v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
              const int r = r_lo + k;
              if (r <= r_hi) {
                const float* src = src01 + r * a.win_stride - s_st[2][r];
                if (VEC == 4) v[k] = __ldcs(reinterpret_cast<const float4*>(src));
                else v[k].x = __ldcs(src);
              }
            }
#pragma unroll
            for (int k = 0; k < 3; ++k)
              if (r_lo + k <= r_hi) {
                acc[0] = __fadd_rn(acc[0], v[k].x);
                if (VEC == 4) acc[1] = __fadd_rn(acc[1], v[k].y), acc[2] = __fadd_rn(acc[2], v[k].z), acc[3] = __fadd_rn(acc[3], v[k].w);
              }
            for (int r = r_lo + 3; r <= r_hi; ++r) {
              const float* src = src01 + r * a.win_stride - s_st[2][r];
              if (VEC == 4) {
                const float4 t = __ldcs(reinterpret_cast<const float4*>(src));
                acc[0] = __fadd_rn(acc[0], t.x), acc[1] = __fadd_rn(acc[1], t.y);
                acc[2] = __fadd_rn(acc[2], t.z), acc[3] = __fadd_rn(acc[3], t.w);
              } else {
                acc[0] = __fadd_rn(acc[0], __ldcs(src));
              }
            }
          }
        }
        if (zin) {
#pragma unroll
          for (int i = 0; i < VEC; ++i) acc[i] = __fdiv_rn(acc[i], s_count[z + i]);
          const long long v = ((long long)xl * a.d1 + y) * a.d2 + z;
          if (VEC == 4) {
            *reinterpret_cast<float4*>(vb + c * CH + lane * 4) = make_float4(acc[0], acc[1], acc[2], acc[3]);
            if (a.logits) __stcs(reinterpret_cast<float4*>(a.logits + c * vox + v), make_float4(acc[0], acc[1], acc[2], acc[3]));
          } else {
            vb[c * CH + lane] = acc[0];
            if (a.logits) __stcs(a.logits + c * vox + v, acc[0]);
          }
        }
      }
      __syncthreads();  // double-buffered: the next chunk writes the other half while this one is reduced
      // ---- argmax (ties -> lowest class, as torch.argmax) / softmax across channels: one voxel per thread
      for (int t = threadIdx.x; t < CH; t += blockDim.x) {
        const int zz = ch * CH + t;
        if (zz >= a.d2) continue;
        float best = vb[t];
        int arg = 0;
        for (int c = 1; c < C; ++c) {
          const float f = vb[c * CH + t];
          if (f > best) best = f, arg = c;
        }
        const long long v = ((long long)xl * a.d1 + y) * a.d2 + zz;
        if (a.labels) a.labels[v] = (uint8_t)arg;
        if (a.probs) {
          float sum = 0.f;
          for (int c = 0; c < C; ++c) sum += expf(vb[c * CH + t] - best);
          const float inv = 1.f / sum;
          for (int c = 0; c < C; ++c) __stcs(a.probs + c * vox + v, expf(vb[c * CH + t] - best) * inv);
        }
      }
    }
    __syncthreads();  // the count map of the row is rewritten next
  }
}

// Streaming form of the deferred blend (labels, optionally logits): one thread owns FOUR consecutive axis-2 voxels and
// ALL classes (CT per pass) -- the covering-window address arithmetic is done once per window instead of once per
// (window, class), the class reduction needs no shared memory or block barrier, and a thread has CT independent
// 16-byte loads in flight per window (a warp reads CT contiguous 512-byte runs).  ~10 instructions per load instead
// of ~110 in the warp-per-class kernel above, which stays for VEC = 1 geometries and the softmax output.
//
// argmax(acc / count) without dividing every class: x -> fdiv_rn(x, count) is monotonic (count > 0), so the winner is
// the lowest class whose QUOTIENT equals the quotient of the largest sum.  Scanning classes upwards, a sum that is
// larger by more than 2^-22 relative certainly has a larger quotient; anything closer (or tiny magnitudes, where the
// quotient could underflow) is decided by the two exact divisions -- the labels are bit-identical to
// `torch.argmax(acc / count)` while the common case does no division at all.
template <int VEC> struct VecT;
template <> struct VecT<4> { using F = float4; using U = uchar4; };
template <> struct VecT<2> { using F = float2; using U = uchar2; };
__device__ __forceinline__ void unpack_vec(const float4& v, float* x) { x[0] = v.x, x[1] = v.y, x[2] = v.z, x[3] = v.w; }
__device__ __forceinline__ void unpack_vec(const float2& v, float* x) { x[0] = v.x, x[1] = v.y; }
__device__ __forceinline__ void store_vec(float* p, const float* x, float4*) { __stcs(reinterpret_cast<float4*>(p), make_float4(x[0], x[1], x[2], x[3])); }
__device__ __forceinline__ void store_vec(float* p, const float* x, float2*) { __stcs(reinterpret_cast<float2*>(p), make_float2(x[0], x[1])); }
__device__ __forceinline__ void store_lab(uint8_t* p, const int* a, uchar4*) { *reinterpret_cast<uchar4*>(p) = make_uchar4((uint8_t)a[0], (uint8_t)a[1], (uint8_t)a[2], (uint8_t)a[3]); }
__device__ __forceinline__ void store_lab(uint8_t* p, const int* a, uchar2*) { *reinterpret_cast<uchar2*>(p) = make_uchar2((uint8_t)a[0], (uint8_t)a[1]); }

// VEC = 4 (or 2): consecutive axis-2 voxels per thread; needs dims[2], roi[2] and every axis-2 window start to be
// multiples of VEC (aligned vector loads inside every covering window).  E.g. 256 x 256 x 358 (BASELINE configs[2]
// after Spacing): the last axis-2 start is 358 - 96 = 262 -> VEC = 2.
template <int CT, bool FULL, int VEC>
__global__ void __launch_bounds__(128, 4) gather_blend_vt_kernel(const GatherArgs a) {
  using F = typename VecT<VEC>::F;
  using U = typename VecT<VEC>::U;
  __shared__ int s_st[3][SGM_MAX_STARTS];
  __shared__ float s_im[3][512];
  for (int i = threadIdx.x; i < 3 * SGM_MAX_STARTS; i += blockDim.x) s_st[i / SGM_MAX_STARTS][i % SGM_MAX_STARTS] = a.starts[i];
  for (int i = threadIdx.x; i < a.roi[0]; i += blockDim.x) s_im[0][i] = a.imap0[i];
  for (int i = threadIdx.x; i < a.roi[1]; i += blockDim.x) s_im[1][i] = a.imap1[i];
  for (int i = threadIdx.x; i < a.roi[2]; i += blockDim.x) s_im[2][i] = a.imap2[i];
  __syncthreads();
  const int C = a.channels;
  const int nq = a.d2 / VEC;  // voxel groups per row
  const long long vox = (long long)a.nx * a.d1 * a.d2;
  const long long total = (long long)a.nx * a.d1 * nq;
  const long long step = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += step) {
    const int zq = (int)(i % nq);
    const long long row = i / nq;
    const int y = (int)(row % a.d1), xl = (int)(row / a.d1);
    const int x = xl + a.x0, z = zq * VEC;
    int p_lo, p_hi, q_lo, q_hi, r_lo, r_hi;
    cover_range(s_st[0], a.n_starts[0], a.roi[0], x, p_lo, p_hi);
    cover_range(s_st[1], a.n_starts[1], a.roi[1], y, q_lo, q_hi);
    cover_range(s_st[2], a.n_starts[2], a.roi[2], z, r_lo, r_hi);
    // ---- count map, MONAI's window order and roundings
    float cnt[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) cnt[k] = 0.f;
    for (int p = p_lo; p <= p_hi; ++p) {
      const float g0 = s_im[0][x - s_st[0][p]];
      for (int q = q_lo; q <= q_hi; ++q) {
        const float g01 = __fmul_rn(g0, s_im[1][y - s_st[1][q]]);
        for (int r = r_lo; r <= r_hi; ++r) {
          const float* g2 = &s_im[2][z - s_st[2][r]];
#pragma unroll
          for (int k = 0; k < VEC; ++k) cnt[k] = __fadd_rn(cnt[k], fmaxf(__fmul_rn(g01, g2[k]), a.floor));
        }
      }
    }
    const int pa = max(p_lo, a.a0_begin), pb = min(p_hi, a.a0_end - 1);  // windows of this rank
    const long long v = ((long long)xl * a.d1 + y) * a.d2 + z;
    float best[VEC], qbest[VEC];   // undivided sum of the current winner; logits mode: its quotient
    int arg[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) best[k] = 0.f, qbest[k] = 0.f, arg[k] = 0;
    for (int c0 = 0; c0 < C; c0 += CT) {
      float acc[CT][VEC];
#pragma unroll
      for (int c = 0; c < CT; ++c)
#pragma unroll
        for (int k = 0; k < VEC; ++k) acc[c][k] = 0.f;
      const float* wl_c = a.wl + (long long)c0 * a.cstride + z;
      for (int p = pa; p <= pb; ++p) {
        const long long w0 = (long long)(p - a.a0_begin) * a.n_starts[1];
        const int off0 = (x - s_st[0][p]) * a.roi[1];
        for (int q = q_lo; q <= q_hi; ++q) {
          const float* src01 = wl_c + (w0 + q) * a.n_starts[2] * a.win_stride + (long long)(off0 + (y - s_st[1][q])) * a.roi[2];
          for (int r = r_lo; r <= r_hi; ++r) {
            const float* src = src01 + r * a.win_stride - s_st[2][r];
            F t[CT];
#pragma unroll
            for (int c = 0; c < CT; ++c)
              if (FULL || c0 + c < C) t[c] = __ldcs(reinterpret_cast<const F*>(src + c * a.cstride));
#pragma unroll
            for (int c = 0; c < CT; ++c)
              if (FULL || c0 + c < C) {
                float tv[VEC];
                unpack_vec(t[c], tv);
#pragma unroll
                for (int k = 0; k < VEC; ++k) acc[c][k] = __fadd_rn(acc[c][k], tv[k]);
              }
          }
        }
      }
      if (a.logits) {  // normalised logits wanted: divide everything, plain argmax on the quotients
#pragma unroll
        for (int c = 0; c < CT; ++c)
          if (FULL || c0 + c < C) {
            float qv[VEC];
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
              qv[k] = __fdiv_rn(acc[c][k], cnt[k]);
              if (c0 + c == 0 || qv[k] > qbest[k]) qbest[k] = qv[k], arg[k] = c0 + c;
            }
            store_vec(a.logits + (long long)(c0 + c) * vox + v, qv, (F*)nullptr);
          }
      } else {
#pragma unroll
        for (int c = 0; c < CT; ++c)
          if (FULL || c0 + c < C) {
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
              const float s = acc[c][k], m = best[k];
              if (c0 + c == 0) {
                best[k] = s;
              } else if (s > m) {
                const float mag = fmaxf(fabsf(s), fabsf(m));
                const bool clear = (s - m) > mag * 2.384185791015625e-07f && fminf(fabsf(s), fabsf(m)) > 1e-30f && mag < 1e30f;
                if (clear || __fdiv_rn(s, cnt[k]) > __fdiv_rn(m, cnt[k])) best[k] = s, arg[k] = c0 + c;
              }
            }
          }
      }
    }
    if (a.labels) store_lab(a.labels + v, arg, (U*)nullptr);
  }
}

template <int CT, int VEC>
int launch_gather_vt(const GatherArgs& a, cudaStream_t st) {
  const long long total = (long long)a.nx * a.d1 * (a.d2 / VEC);
  const int blocks = (int)std::max<long long>(1, std::min<long long>((total + 127) / 128, 148LL * 32));
  if (a.channels % CT == 0)
    gather_blend_vt_kernel<CT, true, VEC><<<blocks, 128, 0, st>>>(a);
  else
    gather_blend_vt_kernel<CT, false, VEC><<<blocks, 128, 0, st>>>(a);
  SGM_CUDA_CHECK(cudaGetLastError());
  return SGM_OK;
}

template <int VEC>
int launch_gather_vt_c(const GatherArgs& a, int channels, cudaStream_t st) {
  // classes per pass: a divisor of C keeps the unrolled class loops free of guards (no spills)
  if (channels <= 4) return launch_gather_vt<4, VEC>(a, st);
  if (channels % 10 == 0) return launch_gather_vt<10, VEC>(a, st);
  if (channels <= 8 || channels % 8 == 0) return launch_gather_vt<8, VEC>(a, st);
  return launch_gather_vt<10, VEC>(a, st);
}

}  // namespace

int launch_gather_blend(const float* wl, int channels, const sgm_sw_cfg* cfg, const int* starts_dev,
                        const float* imap_dev[3], float* logits, uint8_t* labels, float* probs, cudaStream_t st) {
  GatherArgs a;
  a.wl = wl;
  a.cstride = (long long)cfg->roi[0] * cfg->roi[1] * cfg->roi[2];
  a.win_stride = a.cstride * channels;
  a.logits = logits, a.labels = labels, a.probs = probs, a.channels = channels;
  a.nx = cfg->acc_nx, a.d1 = cfg->dims[1], a.d2 = cfg->dims[2], a.x0 = cfg->acc_x0;
  for (int i = 0; i < 3; ++i) a.roi[i] = cfg->roi[i], a.n_starts[i] = cfg->n_starts[i];
  a.a0_begin = cfg->a0_begin, a.a0_end = cfg->a0_end;
  a.starts = starts_dev;
  a.imap0 = imap_dev[0], a.imap1 = imap_dev[1], a.imap2 = imap_dev[2];
  a.floor = cfg->imap_floor;
  auto aligned = [&](int m) {
    bool ok = (a.d2 % m == 0) && (a.roi[2] % m == 0);
    for (int j = 0; j < cfg->n_starts[2]; ++j) ok = ok && (cfg->starts[2][j] % m == 0);
    return ok;
  };
  const bool vec4 = aligned(4), vec2 = aligned(2);
  if (channels > 64) {
    set_error("deferred blend supports at most 64 classes, got %d", channels);
    return SGM_ERR_UNSUPPORTED;
  }
  static const bool force_cw = getenv("SGM_BLEND_CW") != nullptr;  // A/B switch: the warp-per-class kernel
  if ((vec4 || vec2) && !probs && !force_cw && std::max(a.roi[0], std::max(a.roi[1], a.roi[2])) <= 512)
    return vec4 ? launch_gather_vt_c<4>(a, channels, st) : launch_gather_vt_c<2>(a, channels, st);
  const int nwarps = std::min(channels, kGatherWarps);
  const long long nrows = (long long)a.nx * a.d1;
  const int blocks = (int)std::max<long long>(1, std::min<long long>(nrows, 148LL * 64));
  const size_t smem = ((size_t)2 * channels * 32 * (vec4 ? 4 : 1) + (size_t)a.d2 + 8) * sizeof(float);
  SGM_REQUIRE(smem <= 96 * 1024, SGM_ERR_UNSUPPORTED, "deferred blend: axis 2 of %d voxels does not fit shared memory", a.d2);
  static DeviceOnce attr_once;
  if (attr_once.first()) {
    SGM_CUDA_CHECK(cudaFuncSetAttribute(gather_blend_cw_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    SGM_CUDA_CHECK(cudaFuncSetAttribute(gather_blend_cw_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
  }
  if (vec4) {
    gather_blend_cw_kernel<4><<<blocks, nwarps * 32, smem, st>>>(a, nwarps);
  } else {
    gather_blend_cw_kernel<1><<<blocks, nwarps * 32, smem, st>>>(a, nwarps);
  }
  SGM_CUDA_CHECK(cudaGetLastError());
  return SGM_OK;
}

int launch_finalize(const float* acc, int channels, const sgm_sw_cfg* cfg, const int* starts_dev,
                    const float* imap_dev[3], float* logits, uint8_t* labels, float* probs,
                    cudaStream_t st) {
  FinalizeArgs a;
  a.acc = acc, a.logits = logits, a.labels = labels, a.probs = probs;
  a.channels = channels;
  a.nx = cfg->acc_nx, a.d1 = cfg->dims[1], a.d2 = cfg->dims[2], a.x0 = cfg->acc_x0;
  for (int i = 0; i < 3; ++i) a.roi[i] = cfg->roi[i], a.n_starts[i] = cfg->n_starts[i];
  a.starts = starts_dev;
  a.imap0 = imap_dev[0], a.imap1 = imap_dev[1], a.imap2 = imap_dev[2];
  a.floor = cfg->imap_floor;
  const long long vox = (long long)a.nx * a.d1 * a.d2;
  int blocks = (int)((vox + 255) / 256);
  const int cap = 148 * 8 * 4;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  if (channels <= 4)
    finalize_kernel<4><<<blocks, 256, 0, st>>>(a);
  else if (channels <= 8)
    finalize_kernel<8><<<blocks, 256, 0, st>>>(a);
  else if (channels <= 16)
    finalize_kernel<16><<<blocks, 256, 0, st>>>(a);
  else if (channels <= 32)
    finalize_kernel<32><<<blocks, 256, 0, st>>>(a);
  else if (channels <= 64)
    finalize_kernel<64><<<blocks, 256, 0, st>>>(a);
  else {
    set_error("sgm_sw_finalize supports at most 64 classes, got %d", channels);
    return SGM_ERR_UNSUPPORTED;
  }
  SGM_CUDA_CHECK(cudaGetLastError());
  return SGM_OK;
}

}  // namespace sgm
