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

namespace sgm {

namespace {

constexpr int MAX_COVER = 8;

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

__device__ __forceinline__ int covering(const int* starts, int ns, int roi, int v, int* idx) {
  int c = 0;
  for (int j = 0; j < ns; ++j) {
    const int s = starts[j];
    if (s <= v && v < s + roi && c < MAX_COVER) idx[c++] = v - s;
  }
  return c;
}

template <int CMAX>
__global__ void __launch_bounds__(256) finalize_kernel(const FinalizeArgs a) {
  const long long vox = (long long)a.nx * a.d1 * a.d2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < vox; v += stride) {
    const int z = (int)(v % a.d2);
    const long long t = v / a.d2;
    const int y = (int)(t % a.d1);
    const int x = (int)(t / a.d1) + a.x0;
    int i0[MAX_COVER], i1[MAX_COVER], i2[MAX_COVER];
    const int c0 = covering(a.starts, a.n_starts[0], a.roi[0], x, i0);
    const int c1 = covering(a.starts + SGM_MAX_STARTS, a.n_starts[1], a.roi[1], y, i1);
    const int c2 = covering(a.starts + 2 * SGM_MAX_STARTS, a.n_starts[2], a.roi[2], z, i2);
    float count = 0.f;
    for (int p = 0; p < c0; ++p) {
      const float g0 = a.imap0[i0[p]];
      for (int q = 0; q < c1; ++q) {
        const float g01 = __fmul_rn(g0, a.imap1[i1[q]]);
        for (int r = 0; r < c2; ++r)
          count = __fadd_rn(count, fmaxf(__fmul_rn(g01, a.imap2[i2[r]]), a.floor));
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

__device__ __forceinline__ int covering2(const int* starts, int ns, int roi, int v, int* loc, int* idx) {
  int c = 0;
  for (int j = 0; j < ns; ++j) {
    const int s = starts[j];
    if (s <= v && v < s + roi && c < MAX_COVER) {
      loc[c] = v - s;
      idx[c] = j;
      ++c;
    }
  }
  return c;
}

template <int CMAX>
__global__ void __launch_bounds__(256) gather_blend_kernel(const GatherArgs a) {
  const long long vox = (long long)a.nx * a.d1 * a.d2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < vox; v += stride) {
    const int z = (int)(v % a.d2);
    const long long t = v / a.d2;
    const int y = (int)(t % a.d1);
    const int x = (int)(t / a.d1) + a.x0;
    int l0[MAX_COVER], l1[MAX_COVER], l2[MAX_COVER], j0[MAX_COVER], j1[MAX_COVER], j2[MAX_COVER];
    const int c0 = covering2(a.starts, a.n_starts[0], a.roi[0], x, l0, j0);
    const int c1 = covering2(a.starts + SGM_MAX_STARTS, a.n_starts[1], a.roi[1], y, l1, j1);
    const int c2 = covering2(a.starts + 2 * SGM_MAX_STARTS, a.n_starts[2], a.roi[2], z, l2, j2);
    float acc[CMAX];
#pragma unroll
    for (int c = 0; c < CMAX; ++c) acc[c] = 0.f;
    float count = 0.f;
    for (int p = 0; p < c0; ++p) {
      const float g0 = a.imap0[l0[p]];
      const bool have = j0[p] >= a.a0_begin && j0[p] < a.a0_end;
      for (int q = 0; q < c1; ++q) {
        const float g01 = __fmul_rn(g0, a.imap1[l1[q]]);
        for (int r = 0; r < c2; ++r) {
          count = __fadd_rn(count, fmaxf(__fmul_rn(g01, a.imap2[l2[r]]), a.floor));
          if (!have) continue;
          const long long w = ((long long)(j0[p] - a.a0_begin) * a.n_starts[1] + j1[q]) * a.n_starts[2] + j2[r];
          const float* src = a.wl + w * a.win_stride + ((long long)l0[p] * a.roi[1] + l1[q]) * a.roi[2] + l2[r];
#pragma unroll
          for (int c = 0; c < CMAX; ++c)
            if (c < a.channels) acc[c] = __fadd_rn(acc[c], __ldcs(src + c * a.cstride));
        }
      }
    }
    float best = 0.f;
    int arg = 0;
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
      if (c < a.channels) {
        acc[c] = __fdiv_rn(acc[c], count);
        if (c == 0 || acc[c] > best) best = acc[c], arg = c;
      }
    }
    if (a.labels) a.labels[v] = (uint8_t)arg;
    if (a.logits) {
#pragma unroll
      for (int c = 0; c < CMAX; ++c)
        if (c < a.channels) __stcs(a.logits + c * vox + v, acc[c]);
    }
    if (a.probs) {
      float sum = 0.f;
#pragma unroll
      for (int c = 0; c < CMAX; ++c)
        if (c < a.channels) {
          acc[c] = expf(acc[c] - best);
          sum += acc[c];
        }
      const float inv = 1.f / sum;
#pragma unroll
      for (int c = 0; c < CMAX; ++c)
        if (c < a.channels) __stcs(a.probs + c * vox + v, acc[c] * inv);
    }
  }
}

// Vectorised form: one thread owns 4 consecutive voxels along the fastest axis and moves 16 bytes per
// (window, channel) -- the scalar form is bound by load-instruction issue (one 4-byte LDG per lane per
// window-channel), not by HBM.  Needs roi[2], dims[2] and every axis-2 window start to be multiples of 4
// (then the 4 voxels share their covering windows and every float4 is aligned).
template <int CMAX>
__global__ void __launch_bounds__(256) gather_blend_kernel_v4(const GatherArgs a) {
  const int d2q = a.d2 >> 2;
  const long long vox = (long long)a.nx * a.d1 * a.d2;
  const long long nq = (long long)a.nx * a.d1 * d2q;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long qi = (long long)blockIdx.x * blockDim.x + threadIdx.x; qi < nq; qi += stride) {
    const int zq = (int)(qi % d2q);
    const long long t = qi / d2q;
    const int y = (int)(t % a.d1);
    const int xl = (int)(t / a.d1);
    const int x = xl + a.x0, z = zq * 4;
    int l0[MAX_COVER], l1[MAX_COVER], l2[MAX_COVER], j0[MAX_COVER], j1[MAX_COVER], j2[MAX_COVER];
    const int c0 = covering2(a.starts, a.n_starts[0], a.roi[0], x, l0, j0);
    const int c1 = covering2(a.starts + SGM_MAX_STARTS, a.n_starts[1], a.roi[1], y, l1, j1);
    const int c2 = covering2(a.starts + 2 * SGM_MAX_STARTS, a.n_starts[2], a.roi[2], z, l2, j2);
    float acc[CMAX][4];
#pragma unroll
    for (int c = 0; c < CMAX; ++c) acc[c][0] = acc[c][1] = acc[c][2] = acc[c][3] = 0.f;
    float count[4] = {0.f, 0.f, 0.f, 0.f};
    for (int p = 0; p < c0; ++p) {
      const float g0 = a.imap0[l0[p]];
      const bool have = j0[p] >= a.a0_begin && j0[p] < a.a0_end;
      for (int q = 0; q < c1; ++q) {
        const float g01 = __fmul_rn(g0, a.imap1[l1[q]]);
        for (int r = 0; r < c2; ++r) {
          const float4 g2 = *reinterpret_cast<const float4*>(a.imap2 + l2[r]);
          count[0] = __fadd_rn(count[0], fmaxf(__fmul_rn(g01, g2.x), a.floor));
          count[1] = __fadd_rn(count[1], fmaxf(__fmul_rn(g01, g2.y), a.floor));
          count[2] = __fadd_rn(count[2], fmaxf(__fmul_rn(g01, g2.z), a.floor));
          count[3] = __fadd_rn(count[3], fmaxf(__fmul_rn(g01, g2.w), a.floor));
          if (!have) continue;
          const long long w = ((long long)(j0[p] - a.a0_begin) * a.n_starts[1] + j1[q]) * a.n_starts[2] + j2[r];
          const float* src = a.wl + w * a.win_stride + ((long long)l0[p] * a.roi[1] + l1[q]) * a.roi[2] + l2[r];
          float4 x4[CMAX];
#pragma unroll
          for (int c = 0; c < CMAX; ++c)
            if (c < a.channels) x4[c] = __ldcs(reinterpret_cast<const float4*>(src + c * a.cstride));
#pragma unroll
          for (int c = 0; c < CMAX; ++c)
            if (c < a.channels) {
              acc[c][0] = __fadd_rn(acc[c][0], x4[c].x);
              acc[c][1] = __fadd_rn(acc[c][1], x4[c].y);
              acc[c][2] = __fadd_rn(acc[c][2], x4[c].z);
              acc[c][3] = __fadd_rn(acc[c][3], x4[c].w);
            }
        }
      }
    }
    const long long v = ((long long)xl * a.d1 + y) * a.d2 + z;
    float best[4];
    int arg[4] = {0, 0, 0, 0};
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
      if (c < a.channels) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          acc[c][i] = __fdiv_rn(acc[c][i], count[i]);
          if (c == 0 || acc[c][i] > best[i]) best[i] = acc[c][i], arg[i] = c;
        }
      }
    }
    if (a.labels)
      *reinterpret_cast<uchar4*>(a.labels + v) = make_uchar4((uint8_t)arg[0], (uint8_t)arg[1], (uint8_t)arg[2], (uint8_t)arg[3]);
    if (a.logits) {
#pragma unroll
      for (int c = 0; c < CMAX; ++c)
        if (c < a.channels)
          __stcs(reinterpret_cast<float4*>(a.logits + c * vox + v), make_float4(acc[c][0], acc[c][1], acc[c][2], acc[c][3]));
    }
    if (a.probs) {
      float sum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int c = 0; c < CMAX; ++c)
        if (c < a.channels) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            acc[c][i] = expf(acc[c][i] - best[i]);
            sum[i] += acc[c][i];
          }
        }
#pragma unroll
      for (int c = 0; c < CMAX; ++c)
        if (c < a.channels)
          __stcs(reinterpret_cast<float4*>(a.probs + c * vox + v),
                 make_float4(acc[c][0] * (1.f / sum[0]), acc[c][1] * (1.f / sum[1]), acc[c][2] * (1.f / sum[2]),
                             acc[c][3] * (1.f / sum[3])));
    }
  }
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
  const long long vox = (long long)a.nx * a.d1 * a.d2;
  bool vec4 = (a.d2 % 4 == 0) && (a.roi[2] % 4 == 0) && channels <= 16;
  for (int j = 0; j < cfg->n_starts[2]; ++j) vec4 = vec4 && (cfg->starts[2][j] % 4 == 0);
  if (vec4) {
    const int blocks4 = (int)std::max<long long>(1, std::min<long long>((vox / 4 + 255) / 256, 148LL * 8 * 8));
    if (channels <= 4)
      gather_blend_kernel_v4<4><<<blocks4, 256, 0, st>>>(a);
    else if (channels <= 8)
      gather_blend_kernel_v4<8><<<blocks4, 256, 0, st>>>(a);
    else
      gather_blend_kernel_v4<16><<<blocks4, 256, 0, st>>>(a);
    SGM_CUDA_CHECK(cudaGetLastError());
    return SGM_OK;
  }
  int blocks = (int)std::min<long long>((vox + 255) / 256, 148LL * 8 * 8);
  if (blocks < 1) blocks = 1;
  if (channels <= 4)
    gather_blend_kernel<4><<<blocks, 256, 0, st>>>(a);
  else if (channels <= 8)
    gather_blend_kernel<8><<<blocks, 256, 0, st>>>(a);
  else if (channels <= 16)
    gather_blend_kernel<16><<<blocks, 256, 0, st>>>(a);
  else if (channels <= 32)
    gather_blend_kernel<32><<<blocks, 256, 0, st>>>(a);
  else if (channels <= 64)
    gather_blend_kernel<64><<<blocks, 256, 0, st>>>(a);
  else {
    set_error("deferred blend supports at most 64 classes, got %d", channels);
    return SGM_ERR_UNSUPPORTED;
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
