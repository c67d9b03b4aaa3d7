// Bandwidth-bound resamplers and the small pre-processing reductions of the prediction path.
//
//   resample_trilinear[_argmax]  MONAI Spacingd / Invertd(Spacingd): F.grid_sample(bilinear, border,
//                                align_corners=False) on index coordinates, float64 arithmetic, the
//                                eight corners summed in ATen's order (seg/monai_unet.py:173-174,615-621).
//   resample_itk                 sitk.ResampleImageFilter, identity transform, nearest | linear
//                                (image/processing.py:60-70,87-97), ITK's scan-line index formula.
//   normalize_intensity, foreground_bbox   NormalizeIntensityd / CropForegroundd (monai_unet.py:163-169).
//
// One thread per output voxel, lanes along the fastest axis so stores (and, for near-axis-aligned
// grids, loads) coalesce; inputs are read through the read-only path and stay L2-resident (a 2:1
// resample touches each input line from 8 neighbouring outputs).
#include "common.cuh"

#include <limits.h>
#include <math.h>
#include <stdlib.h>

#include <algorithm>

namespace sgm {

namespace {

struct TriArgs {
  const float* in;
  float* out;
  uint8_t* labels;
  int channels;
  int id0, id1, id2;
  int od0, od1, od2;
  double m[12];
};

__device__ __forceinline__ double clipd(double x, int n) { return fmin((double)(n - 1), fmax(x, 0.0)); }

template <bool ARGMAX>
__global__ void __launch_bounds__(256) trilinear_kernel(const TriArgs a) {
  const long long ovox = (long long)a.od0 * a.od1 * a.od2;
  const long long ivox = (long long)a.id0 * a.id1 * a.id2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < ovox; v += stride) {
    const int o2 = (int)(v % a.od2);
    const long long t = v / a.od2;
    const int o1 = (int)(t % a.od1);
    const int o0 = (int)(t / a.od1);
    // source index = A * o + t  (axis 0 = "z" of grid_sample, axis 2 = "x")
    double c0 = a.m[0] * o0 + a.m[1] * o1 + a.m[2] * o2 + a.m[3];
    double c1 = a.m[4] * o0 + a.m[5] * o1 + a.m[6] * o2 + a.m[7];
    double c2 = a.m[8] * o0 + a.m[9] * o1 + a.m[10] * o2 + a.m[11];
    c0 = clipd(c0, a.id0), c1 = clipd(c1, a.id1), c2 = clipd(c2, a.id2);
    const double f0 = floor(c0), f1 = floor(c1), f2 = floor(c2);
    const int z0 = (int)f0, y0 = (int)f1, x0 = (int)f2;
    const int z1 = z0 + 1, y1 = y0 + 1, x1 = x0 + 1;
    const double wx1 = c2 - f2, wx0 = (f2 + 1.0) - c2;
    const double wy1 = c1 - f1, wy0 = (f1 + 1.0) - c1;
    const double wz1 = c0 - f0, wz0 = (f0 + 1.0) - c0;
    // ATen order: tnw tne tsw tse bnw bne bsw bse  (t/b = z0/z1, n/s = y0/y1, w/e = x0/x1)
    const double w[8] = {__dmul_rn(__dmul_rn(wx0, wy0), wz0), __dmul_rn(__dmul_rn(wx1, wy0), wz0),
                         __dmul_rn(__dmul_rn(wx0, wy1), wz0), __dmul_rn(__dmul_rn(wx1, wy1), wz0),
                         __dmul_rn(__dmul_rn(wx0, wy0), wz1), __dmul_rn(__dmul_rn(wx1, wy0), wz1),
                         __dmul_rn(__dmul_rn(wx0, wy1), wz1), __dmul_rn(__dmul_rn(wx1, wy1), wz1)};
    const bool bx1 = x1 < a.id2, by1 = y1 < a.id1, bz1 = z1 < a.id0;
    const bool ok[8] = {true, bx1, by1, bx1 && by1, bz1, bz1 && bx1, bz1 && by1, bz1 && by1 && bx1};
    long long off[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int zz = (k & 4) ? z1 : z0, yy = (k & 2) ? y1 : y0, xx = (k & 1) ? x1 : x0;
      off[k] = ok[k] ? ((long long)zz * a.id1 + yy) * a.id2 + xx : 0;
    }
    float best = 0.f;
    int arg = 0;
    for (int c = 0; c < a.channels; ++c) {
      const float* src = a.in + c * ivox;
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (ok[k]) s = __dadd_rn(s, __dmul_rn((double)__ldg(src + off[k]), w[k]));
      const float r = (float)s;
      if (ARGMAX) {
        if (c == 0 || r > best) best = r, arg = c;
      } else {
        a.out[c * ovox + v] = r;
      }
    }
    if (ARGMAX) a.labels[v] = (uint8_t)arg;
  }
}

// Shared-memory-staged form for axis-aligned transforms (Spacingd and its inverse are diagonal scalings + offsets).
// The gather above reads every input line from up to 8 neighbouring outputs that sit in different CTAs; with a 3:1
// ratio along the fastest axis a warp load touches 3-4 lines for 128 useful bytes, and the 10-channel inverse moved
// ~30 GB through L2 for 0.94 GB of input (6.6 ms).  Here a CTA owns a T0 x T1 x T2 tile of outputs (one thread per
// output voxel), stages the source box of CC channels with coalesced row loads, and interpolates from shared memory:
// each input voxel leaves L2 ~1.5-3x instead of ~30x.  The per-voxel arithmetic is the expression sequence of
// trilinear_kernel, statement for statement (float64, ATen corner order): results are bit-identical
// (tests/test_gpu_resample.py compares the two kernels for equality).
struct BrickPlan {
  int T0, T1, T2;     // output tile (threads per CTA = T0 * T1 * T2)
  int S0, S1, S2;     // capacity of the source box along each axis
  int nt0, nt1, nt2;  // tiles per axis
  int CC;             // channels staged per pass
};

template <bool ARGMAX>
__global__ void __launch_bounds__(256, 4) trilinear_brick_kernel(const TriArgs a, const BrickPlan p) {
  extern __shared__ float box[];  // [CC][S0][S1][S2]
  // per-axis tables of the tile: an axis-aligned map is separable, so the clipped coordinate, its floor and the two
  // weights of output index o_a depend on o_a alone -- T0 + T1 + T2 threads compute them once per tile instead of every
  // voxel recomputing all three (the terms of the other axes multiply exact zeros: the same rounded value, bit for bit)
  __shared__ int s_idx_[288];                  // axis 0 at [0, 16), axis 1 at [16, 32), axis 2 at [32, 288)
  __shared__ double s_w0_[288], s_w1_[288];
  int* const s_idx[3] = {s_idx_, s_idx_ + 16, s_idx_ + 32};
  double* const s_w0[3] = {s_w0_, s_w0_ + 16, s_w0_ + 32};
  double* const s_w1[3] = {s_w1_, s_w1_ + 16, s_w1_ + 32};
  const int tid = threadIdx.x, nthr = blockDim.x;
  const long long ovox = (long long)a.od0 * a.od1 * a.od2;
  const long long ivox = (long long)a.id0 * a.id1 * a.id2;
  const int t2 = tid % p.T2, t1 = (tid / p.T2) % p.T1, t0 = tid / (p.T2 * p.T1);
  const long long ntiles = (long long)p.nt0 * p.nt1 * p.nt2;
  const int box_vox = p.S0 * p.S1 * p.S2;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int b2 = (int)(tile % p.nt2), b1 = (int)((tile / p.nt2) % p.nt1), b0 = (int)(tile / ((long long)p.nt2 * p.nt1));
    const int O0 = b0 * p.T0, O1 = b1 * p.T1, O2 = b2 * p.T2;
    const int e0 = min(p.T0, a.od0 - O0), e1 = min(p.T1, a.od1 - O1), e2 = min(p.T2, a.od2 - O2);
    if (tid < p.T0 + p.T1 + p.T2) {
      const int ax = tid < p.T0 ? 0 : (tid < p.T0 + p.T1 ? 1 : 2);
      const int t = tid - (ax == 0 ? 0 : (ax == 1 ? p.T0 : p.T0 + p.T1));
      // the statements of trilinear_kernel with the other two output indices at 0 (their coefficients are 0.0)
      const int o0 = ax == 0 ? O0 + t : 0, o1 = ax == 1 ? O1 + t : 0, o2 = ax == 2 ? O2 + t : 0;
      double c;
      if (ax == 0) c = a.m[0] * o0 + a.m[1] * o1 + a.m[2] * o2 + a.m[3];
      else if (ax == 1) c = a.m[4] * o0 + a.m[5] * o1 + a.m[6] * o2 + a.m[7];
      else c = a.m[8] * o0 + a.m[9] * o1 + a.m[10] * o2 + a.m[11];
      c = clipd(c, ax == 0 ? a.id0 : (ax == 1 ? a.id1 : a.id2));
      const double f = floor(c);
      s_idx[ax][t] = (int)f;
      s_w1[ax][t] = c - f;
      s_w0[ax][t] = (f + 1.0) - c;
    }
    __syncthreads();
    // source box: the map is monotonic along every axis, so the extreme indices sit at the ends of the tables
    const int l0 = min(s_idx[0][0], s_idx[0][e0 - 1]), l1 = min(s_idx[1][0], s_idx[1][e1 - 1]), l2 = min(s_idx[2][0], s_idx[2][e2 - 1]);
    const int n0 = min(max(s_idx[0][0], s_idx[0][e0 - 1]) + 1, a.id0 - 1) - l0 + 1;
    const int n1 = min(max(s_idx[1][0], s_idx[1][e1 - 1]) + 1, a.id1 - 1) - l1 + 1;
    const int n2 = min(max(s_idx[2][0], s_idx[2][e2 - 1]) + 1, a.id2 - 1) - l2 + 1;
    const bool live = t0 < e0 && t1 < e1 && t2 < e2;
    const int o0 = O0 + t0, o1 = O1 + t1, o2 = O2 + t2;
    const long long v = ((long long)o0 * a.od1 + o1) * a.od2 + o2;
    const int u0 = live ? t0 : 0, u1 = live ? t1 : 0, u2 = live ? t2 : 0;
    const int z0 = s_idx[0][u0], y0 = s_idx[1][u1], x0 = s_idx[2][u2];
    const int z1 = z0 + 1, y1 = y0 + 1, x1 = x0 + 1;
    const double wz0 = s_w0[0][u0], wz1 = s_w1[0][u0], wy0 = s_w0[1][u1], wy1 = s_w1[1][u1], wx0 = s_w0[2][u2], wx1 = s_w1[2][u2];
    // ATen order: tnw tne tsw tse bnw bne bsw bse  (t/b = z0/z1, n/s = y0/y1, w/e = x0/x1)
    const double w[8] = {__dmul_rn(__dmul_rn(wx0, wy0), wz0), __dmul_rn(__dmul_rn(wx1, wy0), wz0),
                         __dmul_rn(__dmul_rn(wx0, wy1), wz0), __dmul_rn(__dmul_rn(wx1, wy1), wz0),
                         __dmul_rn(__dmul_rn(wx0, wy0), wz1), __dmul_rn(__dmul_rn(wx1, wy0), wz1),
                         __dmul_rn(__dmul_rn(wx0, wy1), wz1), __dmul_rn(__dmul_rn(wx1, wy1), wz1)};
    const bool bx1 = x1 < a.id2, by1 = y1 < a.id1, bz1 = z1 < a.id0;
    const bool ok[8] = {true, bx1, by1, bx1 && by1, bz1, bz1 && bx1, bz1 && by1, bz1 && by1 && bx1};
    int off[8];
    bool inbox = n0 <= p.S0 && n1 <= p.S1 && n2 <= p.S2;  // always, by the plan's box capacity; checked
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int zz = ((k & 4) ? z1 : z0) - l0, yy = ((k & 2) ? y1 : y0) - l1, xx = ((k & 1) ? x1 : x0) - l2;
      off[k] = ok[k] ? (zz * p.S1 + yy) * p.S2 + xx : 0;
    }
    float best = 0.f;
    int arg = 0;
    for (int cb = 0; cb < a.channels; cb += p.CC) {
      const int cc = min(p.CC, a.channels - cb);
      __syncthreads();  // the previous pass (or tile) no longer reads the box
      // ---- stage [cc][n0][n1][n2]: a warp takes four source rows at a time (four independent row loads in flight per
      // lane), lanes along the fastest source axis
      if (inbox) {
        // a warp takes one (channel, z) slab of n1 rows at a time, four rows in flight per lane, lanes along x: one
        // integer division per slab (the first version decomposed a flat row index with four divisions per row and
        // spent 8x more instructions on staging than on interpolating)
        const int lane = tid & 31, nwarps = nthr >> 5;
        for (int cz = tid >> 5; cz < cc * n0; cz += nwarps) {
          const int c = cz / n0, z = cz - c * n0;
          const float* srow = a.in + (long long)(cb + c) * ivox + ((long long)(l0 + z) * a.id1 + l1) * a.id2 + l2;
          float* drow = box + (size_t)c * box_vox + (size_t)z * p.S1 * p.S2;
          for (int y = 0; y < n1; y += 4) {
            for (int x = lane; x < n2; x += 32) {
              float t[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) t[j] = __ldg(srow + (long long)min(y + j, n1 - 1) * a.id2 + x);
#pragma unroll
              for (int j = 0; j < 4; ++j)
                if (y + j < n1) drow[(y + j) * p.S2 + x] = t[j];
            }
          }
        }
      }
      __syncthreads();
      if (live) {
        for (int c = 0; c < cc; ++c) {
          double sacc = 0.0;
          if (inbox) {
            const float* bsrc = box + (size_t)c * box_vox;
#pragma unroll
            for (int k = 0; k < 8; ++k)
              if (ok[k]) sacc = __dadd_rn(sacc, __dmul_rn((double)bsrc[off[k]], w[k]));
          } else {  // never taken when the plan's capacity holds; kept so that a planning slip costs speed, not results
            const float* gsrc = a.in + (long long)(cb + c) * ivox;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const int zz = (k & 4) ? z1 : z0, yy = (k & 2) ? y1 : y0, xx = (k & 1) ? x1 : x0;
              if (ok[k]) sacc = __dadd_rn(sacc, __dmul_rn((double)__ldg(gsrc + ((long long)zz * a.id1 + yy) * a.id2 + xx), w[k]));
            }
          }
          const float r = (float)sacc;
          if (ARGMAX) {
            if (cb + c == 0 || r > best) best = r, arg = cb + c;
          } else {
            a.out[(long long)(cb + c) * ovox + v] = r;
          }
        }
      }
    }
    if (ARGMAX && live) a.labels[v] = (uint8_t)arg;
    __syncthreads();  // the tables are rewritten for the next tile
  }
}

// ---- axis-aligned maps without staging: separable per-axis tables + one warp per output row.
// The gather kernel above spends most of its instructions on two 64-bit integer divisions per voxel and on recomputing
// the three float64 source coordinates, floors and weights that depend on ONE output index each.  For diagonal maps a
// table kernel evaluates those statements once per output index of every axis (the same expressions: bit-identical
// results), and the main kernel walks output rows -- the axis-0 / axis-1 entries are warp-uniform, the axis-2 entries
// come from the L1-resident table, row addressing is 32-bit.
struct SepTables {
  int* idx;        // [od0 | od1 | od2] floor of the clipped source coordinate
  double* w0;      // weight of the lower neighbour
  double* w1;      // weight of the upper neighbour
};

__global__ void trilinear_tables_kernel(const TriArgs a, SepTables t) {
  const int n = a.od0 + a.od1 + a.od2;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int ax = i < a.od0 ? 0 : (i < a.od0 + a.od1 ? 1 : 2);
    const int o = i - (ax == 0 ? 0 : (ax == 1 ? a.od0 : a.od0 + a.od1));
    // the statements of trilinear_kernel with the other two output indices at 0 (their coefficients are 0.0)
    const int o0 = ax == 0 ? o : 0, o1 = ax == 1 ? o : 0, o2 = ax == 2 ? o : 0;
    double c;
    if (ax == 0) c = a.m[0] * o0 + a.m[1] * o1 + a.m[2] * o2 + a.m[3];
    else if (ax == 1) c = a.m[4] * o0 + a.m[5] * o1 + a.m[6] * o2 + a.m[7];
    else c = a.m[8] * o0 + a.m[9] * o1 + a.m[10] * o2 + a.m[11];
    c = clipd(c, ax == 0 ? a.id0 : (ax == 1 ? a.id1 : a.id2));
    const double f = floor(c);
    t.idx[i] = (int)f;
    t.w1[i] = c - f;
    t.w0[i] = (f + 1.0) - c;
  }
}

template <bool ARGMAX>
__global__ void __launch_bounds__(256) trilinear_sep_kernel(const TriArgs a, const SepTables t) {
  const long long ovox = (long long)a.od0 * a.od1 * a.od2;
  const long long ivox = (long long)a.id0 * a.id1 * a.id2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int nrows = a.od0 * a.od1;
  const int* ix2 = t.idx + a.od0 + a.od1;
  const double* w02 = t.w0 + a.od0 + a.od1;
  const double* w12 = t.w1 + a.od0 + a.od1;
  for (int row = blockIdx.x * nwarps + warp; row < nrows; row += gridDim.x * nwarps) {
    const int o0 = row / a.od1, o1 = row - o0 * a.od1;
    const int z0 = __ldg(t.idx + o0), y0 = __ldg(t.idx + a.od0 + o1);
    const int z1 = z0 + 1, y1 = y0 + 1;
    const double wz0 = __ldg(t.w0 + o0), wz1 = __ldg(t.w1 + o0);
    const double wy0 = __ldg(t.w0 + a.od0 + o1), wy1 = __ldg(t.w1 + a.od0 + o1);
    const bool by1 = y1 < a.id1, bz1 = z1 < a.id0;
    // row bases of the four (z, y) source rows (clamped rows are never read: their ok flags are false)
    const int r00 = (z0 * a.id1 + y0) * a.id2;
    const int r01 = by1 ? (z0 * a.id1 + y1) * a.id2 : r00;
    const int r10 = bz1 ? (z1 * a.id1 + y0) * a.id2 : r00;
    const int r11 = (bz1 && by1) ? (z1 * a.id1 + y1) * a.id2 : r00;
    const long long vrow = (long long)row * a.od2;
    for (int o2 = lane; o2 < a.od2; o2 += 32) {
      const int x0 = __ldg(ix2 + o2), x1 = x0 + 1;
      const double wx0 = __ldg(w02 + o2), wx1 = __ldg(w12 + o2);
      const bool bx1 = x1 < a.id2;
      // ATen order: tnw tne tsw tse bnw bne bsw bse  (t/b = z0/z1, n/s = y0/y1, w/e = x0/x1)
      const double w[8] = {__dmul_rn(__dmul_rn(wx0, wy0), wz0), __dmul_rn(__dmul_rn(wx1, wy0), wz0),
                           __dmul_rn(__dmul_rn(wx0, wy1), wz0), __dmul_rn(__dmul_rn(wx1, wy1), wz0),
                           __dmul_rn(__dmul_rn(wx0, wy0), wz1), __dmul_rn(__dmul_rn(wx1, wy0), wz1),
                           __dmul_rn(__dmul_rn(wx0, wy1), wz1), __dmul_rn(__dmul_rn(wx1, wy1), wz1)};
      const bool ok[8] = {true, bx1, by1, bx1 && by1, bz1, bz1 && bx1, bz1 && by1, bz1 && by1 && bx1};
      const int xe = bx1 ? x1 : x0;
      const int off[8] = {r00 + x0, r00 + xe, r01 + x0, r01 + xe, r10 + x0, r10 + xe, r11 + x0, r11 + xe};
      float best = 0.f;
      int arg = 0;
      for (int c = 0; c < a.channels; ++c) {
        const float* src = a.in + c * ivox;
        double sacc = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (ok[k]) sacc = __dadd_rn(sacc, __dmul_rn((double)__ldg(src + off[k]), w[k]));
        const float r = (float)sacc;
        if (ARGMAX) {
          if (c == 0 || r > best) best = r, arg = c;
        } else {
          a.out[c * ovox + vrow + o2] = r;
        }
      }
      if (ARGMAX) a.labels[vrow + o2] = (uint8_t)arg;
    }
  }
}

static bool tri_diagonal(const TriArgs& a) {
  if (a.m[1] != 0.0 || a.m[2] != 0.0 || a.m[4] != 0.0 || a.m[6] != 0.0 || a.m[8] != 0.0 || a.m[9] != 0.0) return false;
  return (long long)a.id0 * a.id1 * a.id2 < (1LL << 31) && (long long)a.od0 * a.od1 < (1LL << 31);
}

template <bool ARGMAX>
static int launch_trilinear_sep(const TriArgs& a, cudaStream_t st) {
  const int n = a.od0 + a.od1 + a.od2;
  char* buf = nullptr;
  SGM_CUDA_CHECK(cudaMallocAsync((void**)&buf, (size_t)n * 24, st));  // doubles first (8-byte aligned), then the ints
  SepTables t;
  t.w0 = (double*)buf, t.w1 = t.w0 + n, t.idx = (int*)(t.w1 + n);
  trilinear_tables_kernel<<<(n + 255) / 256, 256, 0, st>>>(a, t);
  const int nrows = a.od0 * a.od1;
  const int grid = std::max(1, std::min((nrows + 7) / 8, 148 * 8));
  trilinear_sep_kernel<ARGMAX><<<grid, 256, 0, st>>>(a, t);
  cudaError_t e = cudaGetLastError();
  cudaFreeAsync(buf, st);
  SGM_CUDA_CHECK(e);
  return SGM_OK;
}

// ------------------------------------------------------------------------------------------ ITK
struct ItkArgs {
  const void* in;
  void* out;
  int in_n[3], out_n[3];
  double i2p[9], oorg[3], p2i[9], iorg[3];
  int nearest;
  double defval;
};

template <typename T>
__device__ __forceinline__ T itk_cast(double v);
template <>
__device__ __forceinline__ uint8_t itk_cast<uint8_t>(double v) {
  return (uint8_t)trunc(fmin(255.0, fmax(0.0, v)));
}
template <>
__device__ __forceinline__ int16_t itk_cast<int16_t>(double v) {
  return (int16_t)trunc(fmin(32767.0, fmax(-32768.0, v)));
}
template <>
__device__ __forceinline__ uint16_t itk_cast<uint16_t>(double v) {
  return (uint16_t)trunc(fmin(65535.0, fmax(0.0, v)));
}
template <>
__device__ __forceinline__ int32_t itk_cast<int32_t>(double v) {
  return (int32_t)trunc(fmin(2147483647.0, fmax(-2147483648.0, v)));
}
template <>
__device__ __forceinline__ float itk_cast<float>(double v) {
  return (float)fmin(3.4028234663852886e38, fmax(-3.4028234663852886e38, v));
}

// continuous input index of output index (ox, oy, oz): ITK TransformIndexToPhysicalPoint then
// TransformPhysicalPointToContinuousIndex, sums left to right, no FMA contraction.
__device__ __forceinline__ void itk_cindex(const ItkArgs& a, double ox, double oy, double oz, double c[3]) {
  double ph[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    double s = __dmul_rn(a.i2p[r * 3 + 0], ox);
    s = __dadd_rn(s, __dmul_rn(a.i2p[r * 3 + 1], oy));
    s = __dadd_rn(s, __dmul_rn(a.i2p[r * 3 + 2], oz));
    ph[r] = __dadd_rn(__dadd_rn(s, a.oorg[r]), -a.iorg[r]);
  }
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    double s = __dmul_rn(a.p2i[r * 3 + 0], ph[0]);
    s = __dadd_rn(s, __dmul_rn(a.p2i[r * 3 + 1], ph[1]));
    s = __dadd_rn(s, __dmul_rn(a.p2i[r * 3 + 2], ph[2]));
    c[r] = s;
  }
}

// one output voxel of scan line (oy, oz): cs / ce = continuous input index of the line's first voxel and of one past its last
template <typename T>
__device__ __forceinline__ T itk_voxel(const ItkArgs& a, const T* in, const double cs[3], const double ce[3], int ox) {
  double c[3];
  const double alpha = (double)ox / (double)a.out_n[0];
  bool inside = true;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    c[r] = __dadd_rn(cs[r], __dmul_rn(alpha, __dadd_rn(ce[r], -cs[r])));
    inside = inside && (c[r] >= -0.5) && (c[r] < (double)a.in_n[r] - 0.5);
  }
  if (!inside) return itk_cast<T>(a.defval);
  if (a.nearest) {
    int idx[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      int i = (int)floor(c[r] + 0.5);
      idx[r] = min(max(i, 0), a.in_n[r] - 1);
    }
    return in[((long long)idx[2] * a.in_n[1] + idx[1]) * a.in_n[0] + idx[0]];
  }
  int lo[3], hi[3];
  double fr[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    int b = max((int)floor(c[r]), 0);
    b = min(b, a.in_n[r] - 1);
    fr[r] = fmax(c[r] - (double)b, 0.0);
    lo[r] = b;
    hi[r] = min(b + 1, a.in_n[r] - 1);
  }
  double val[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int x = (k & 1) ? hi[0] : lo[0], y = (k & 2) ? hi[1] : lo[1], z = (k & 4) ? hi[2] : lo[2];
    val[k] = (double)in[((long long)z * a.in_n[1] + y) * a.in_n[0] + x];
  }
  // lerp along x, then y, then z:  a + f*(b - a)
  double vx[4];
#pragma unroll
  for (int k = 0; k < 4; ++k)
    vx[k] = __dadd_rn(val[2 * k], __dmul_rn(fr[0], __dadd_rn(val[2 * k + 1], -val[2 * k])));
  const double vy0 = __dadd_rn(vx[0], __dmul_rn(fr[1], __dadd_rn(vx[1], -vx[0])));
  const double vy1 = __dadd_rn(vx[2], __dmul_rn(fr[1], __dadd_rn(vx[3], -vx[2])));
  return itk_cast<T>(__dadd_rn(vy0, __dmul_rn(fr[2], __dadd_rn(vy1, -vy0))));
}

template <typename T>
__global__ void __launch_bounds__(256) itk_resample_kernel(const ItkArgs a) {
  const long long ovox = (long long)a.out_n[0] * a.out_n[1] * a.out_n[2];
  const long long stride = (long long)gridDim.x * blockDim.x;
  const T* in = reinterpret_cast<const T*>(a.in);
  T* out = reinterpret_cast<T*>(a.out);
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < ovox; v += stride) {
    const int ox = (int)(v % a.out_n[0]);
    const long long t = v / a.out_n[0];
    const int oy = (int)(t % a.out_n[1]);
    const int oz = (int)(t / a.out_n[1]);
    double cs[3], ce[3];
    itk_cindex(a, 0.0, (double)oy, (double)oz, cs);
    itk_cindex(a, (double)a.out_n[0], (double)oy, (double)oz, ce);
    out[v] = itk_voxel<T>(a, in, cs, ce, ox);
  }
}

// VP consecutive voxels of a scan line per thread: the two scan-line end points (2 x 15 float64 operations) are computed
// once per VP voxels instead of once per voxel, and the VP results leave as one 4..16-byte store (a uint8 label map
// written one byte per lane is a 32-byte store per warp).  Same per-voxel arithmetic: bit-identical.  Needs the line
// length to be a multiple of VP (the launcher falls back to the kernel above otherwise).
template <typename T, int VP>
__global__ void __launch_bounds__(256) itk_resample_vec_kernel(const ItkArgs a) {
  const int nq = a.out_n[0] / VP;
  const long long total = (long long)nq * a.out_n[1] * a.out_n[2];
  const long long stride = (long long)gridDim.x * blockDim.x;
  const T* in = reinterpret_cast<const T*>(a.in);
  T* out = reinterpret_cast<T*>(a.out);
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += stride) {
    const int xq = (int)(q % nq);
    const long long t = q / nq;
    const int oy = (int)(t % a.out_n[1]);
    const int oz = (int)(t / a.out_n[1]);
    double cs[3], ce[3];
    itk_cindex(a, 0.0, (double)oy, (double)oz, cs);
    itk_cindex(a, (double)a.out_n[0], (double)oy, (double)oz, ce);
    struct alignas(sizeof(T) * VP) Pack { T v[VP]; } r;
#pragma unroll
    for (int k = 0; k < VP; ++k) r.v[k] = itk_voxel<T>(a, in, cs, ce, xq * VP + k);
    *reinterpret_cast<Pack*>(out + (t * a.out_n[0] + (long long)xq * VP)) = r;
  }
}

// ---- axis-aligned nearest neighbour: separable index tables.
// With diagonal index<->physical matrices (identity directions: Spacingd's inverse, resample_to_ref between images of
// the same orientation) the continuous index along input axis r depends on the output index along axis r alone: in the
// scan-line formula  c[r] = cs[r] + alpha * (ce[r] - cs[r])  the other output indices enter through products with exact
// zeros, and for r != 0 the line's end points coincide (ce[r] - cs[r] = +0).  A table kernel evaluates THE SAME device
// functions (itk_cindex, the alpha division, the inside test, floor(c + 0.5)) once per output index of every axis --
// bit-identical source indices by construction -- and the main kernel is a pure gather: a warp-coalesced row of
// table look-ups and byte loads instead of ~45 float64 operations, one float64 division and two 64-bit integer
// divisions per voxel.
__global__ void itk_tables_kernel(const ItkArgs a, int* tab) {
  // tab: [out_n0 | out_n1 | out_n2] source index per output index, -1 = outside the input buffer
  const int n0 = a.out_n[0], n1 = a.out_n[1], n2 = a.out_n[2];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n0 + n1 + n2; i += gridDim.x * blockDim.x) {
    const int r = i < n0 ? 0 : (i < n0 + n1 ? 1 : 2);
    const int o = i - (r == 0 ? 0 : (r == 1 ? n0 : n0 + n1));
    const int ox = r == 0 ? o : 0, oy = r == 1 ? o : 0, oz = r == 2 ? o : 0;
    double cs[3], ce[3];
    itk_cindex(a, 0.0, (double)oy, (double)oz, cs);
    itk_cindex(a, (double)a.out_n[0], (double)oy, (double)oz, ce);
    const double alpha = (double)ox / (double)a.out_n[0];
    const double c = __dadd_rn(cs[r], __dmul_rn(alpha, __dadd_rn(ce[r], -cs[r])));
    const bool inside = (c >= -0.5) && (c < (double)a.in_n[r] - 0.5);
    const int idx = min(max((int)floor(c + 0.5), 0), a.in_n[r] - 1);
    tab[i] = inside ? idx : -1;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) itk_nearest_sep_kernel(const ItkArgs a, const int* __restrict__ tab) {
  extern __shared__ int s_tx[];  // the axis-0 table: shared by every scan line
  const int n0 = a.out_n[0], n1 = a.out_n[1], n2 = a.out_n[2];
  for (int i = threadIdx.x; i < n0; i += blockDim.x) s_tx[i] = tab[i];
  __syncthreads();
  const T* in = reinterpret_cast<const T*>(a.in);
  T* out = reinterpret_cast<T*>(a.out);
  const T defv = itk_cast<T>(a.defval);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int nlines = n1 * n2;
  constexpr int VP = 4;
  struct alignas(sizeof(T) * VP) Pack { T v[VP]; };
  const bool packed = (n0 % VP == 0) && ((uintptr_t)out % (sizeof(T) * VP) == 0);
  for (int line = blockIdx.x * nwarps + warp; line < nlines; line += gridDim.x * nwarps) {  // one warp per scan line
    const int oz = line / n1, oy = line - oz * n1;
    const int iy = __ldg(tab + n0 + oy), iz = __ldg(tab + n0 + n1 + oz);
    T* orow = out + (long long)line * n0;
    const bool row_in = iy >= 0 && iz >= 0;
    const T* irow = in + (row_in ? ((long long)iz * a.in_n[1] + iy) * a.in_n[0] : 0);
    if (packed) {
      for (int xq = lane; xq < n0 / VP; xq += 32) {
        const int4 ix = *reinterpret_cast<const int4*>(s_tx + xq * VP);
        Pack r;
        r.v[0] = (row_in && ix.x >= 0) ? __ldg(irow + ix.x) : defv;
        r.v[1] = (row_in && ix.y >= 0) ? __ldg(irow + ix.y) : defv;
        r.v[2] = (row_in && ix.z >= 0) ? __ldg(irow + ix.z) : defv;
        r.v[3] = (row_in && ix.w >= 0) ? __ldg(irow + ix.w) : defv;
        *reinterpret_cast<Pack*>(orow + xq * VP) = r;
      }
    } else {
      for (int x = lane; x < n0; x += 32) {
        const int ix = s_tx[x];
        orow[x] = (row_in && ix >= 0) ? __ldg(irow + ix) : defv;
      }
    }
  }
}

static bool itk_diagonal(const ItkArgs& a) {
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c)
      if (r != c && (a.i2p[r * 3 + c] != 0.0 || a.p2i[r * 3 + c] != 0.0)) return false;
  return true;
}

template <typename T>
static int launch_itk_sep(const ItkArgs& a, cudaStream_t st) {
  const int n = a.out_n[0] + a.out_n[1] + a.out_n[2];
  int* tab = nullptr;
  SGM_CUDA_CHECK(cudaMallocAsync((void**)&tab, (size_t)n * sizeof(int), st));
  itk_tables_kernel<<<(n + 255) / 256, 256, 0, st>>>(a, tab);
  const int nlines = a.out_n[1] * a.out_n[2];
  const int grid = std::max(1, std::min((nlines + 7) / 8, 148 * 8));
  itk_nearest_sep_kernel<T><<<grid, 256, (size_t)a.out_n[0] * sizeof(int), st>>>(a, tab);
  cudaError_t e = cudaGetLastError();
  cudaFreeAsync(tab, st);
  SGM_CUDA_CHECK(e);
  return SGM_OK;
}

// ------------------------------------------------------------------------ normalize / bbox
constexpr int RED_BLOCKS = 1024;

__global__ void __launch_bounds__(256) sum_kernel(const float* x, long long n, const double* mean_ptr,
                                                  double* partial) {
  // partial[b] = sum over this block's grid-stride slice of (x - mean)^k; k=1 w/o mean, k=2 with.
  __shared__ double sm[256];
  const double mean = mean_ptr ? *mean_ptr : 0.0;
  double s = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const double v = (double)x[i];
    s += mean_ptr ? (v - mean) * (v - mean) : v;
  }
  sm[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sm[0];
}

__global__ void reduce_final_kernel(const double* partial, int nb, long long n, int is_var, double* out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < nb; ++i) s += partial[i];
    s /= (double)n;
    if (is_var) {
      s = sqrt(s);
      if (s == 0.0) s = 1.0;
    }
    *out = s;
  }
}

__global__ void __launch_bounds__(256) normalize_apply_kernel(const float* x, float* y, long long n,
                                                              const double* mean_std) {
  const float mean = (float)mean_std[0], sd = (float)mean_std[1];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    y[i] = __fdiv_rn(__fadd_rn(x[i], -mean), sd);
}

__global__ void bbox_init_kernel(int* bbox) {
  if (threadIdx.x < 3) bbox[threadIdx.x] = INT_MAX;
  if (threadIdx.x >= 3 && threadIdx.x < 6) bbox[threadIdx.x] = -1;
}

__global__ void __launch_bounds__(256) bbox_kernel(const float* x, int channels, int d0, int d1, int d2,
                                                   int* bbox) {
  const long long vox = (long long)d0 * d1 * d2;
  int lo[3] = {INT_MAX, INT_MAX, INT_MAX}, hi[3] = {-1, -1, -1};
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < vox;
       v += (long long)gridDim.x * blockDim.x) {
    bool fg = false;
    for (int c = 0; c < channels; ++c) fg = fg || (x[c * vox + v] > 0.f);
    if (fg) {
      const int i2 = (int)(v % d2);
      const long long t = v / d2;
      const int i1 = (int)(t % d1), i0 = (int)(t / d1);
      lo[0] = min(lo[0], i0), lo[1] = min(lo[1], i1), lo[2] = min(lo[2], i2);
      hi[0] = max(hi[0], i0), hi[1] = max(hi[1], i1), hi[2] = max(hi[2], i2);
    }
  }
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    for (int o = 16; o > 0; o >>= 1) {
      lo[r] = min(lo[r], __shfl_xor_sync(0xffffffffu, lo[r], o));
      hi[r] = max(hi[r], __shfl_xor_sync(0xffffffffu, hi[r], o));
    }
    if ((threadIdx.x & 31) == 0) {
      if (lo[r] != INT_MAX) atomicMin(bbox + r, lo[r]);
      if (hi[r] >= 0) atomicMax(bbox + 3 + r, hi[r]);
    }
  }
}

__global__ void bbox_final_kernel(int* bbox) {
  if (threadIdx.x == 0) {
    if (bbox[3] < 0) {
      for (int i = 0; i < 6; ++i) bbox[i] = 0;
    } else {
      for (int i = 3; i < 6; ++i) bbox[i] += 1;
    }
  }
}

int grid_for(long long n) {
  long long b = (n + 255) / 256;
  const long long cap = 148 * 8 * 2;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

// Tile search for trilinear_brick_kernel: T0*T1*T2 = 256 threads, source box (times the channels staged per
// pass) within 48 KB, minimise input voxels staged per output voxel.
static bool brick_plan(const TriArgs& a, BrickPlan& p) {
  if (getenv("SGM_NO_RESAMPLE_BRICK") || getenv("SGM_RESAMPLE_SEP_ALWAYS")) return false;
  // measured on the configs[2] shapes (bench.py roofline_resample): 10-class inverse 6.7 -> 3.0 ms, but the single-channel
  // forward resample 0.38 -> 0.62 ms (one channel does not amortise the tile's stage -> barrier -> compute chain)
  if (a.channels < 4 && !getenv("SGM_RESAMPLE_BRICK_ALWAYS")) return false;
  if (a.m[1] != 0.0 || a.m[2] != 0.0 || a.m[4] != 0.0 || a.m[6] != 0.0 || a.m[8] != 0.0 || a.m[9] != 0.0) return false;
  if (a.m[0] == 0.0 || a.m[5] == 0.0 || a.m[10] == 0.0) return false;
  const int budget = 48 * 1024;  // four CTAs per SM next to 6 KB of per-axis tables
  const int od[3] = {a.od0, a.od1, a.od2};
  const double sc[3] = {fabs(a.m[0]), fabs(a.m[5]), fabs(a.m[10])};
  static const int c01[] = {1, 2, 4, 8, 16};
  static const int c2[] = {8, 16, 32, 64, 128, 256};
  double best = 1e30;
  for (int T0 : c01)
    for (int T1 : c01)
      for (int T2 : c2) {
        const int thr = T0 * T1 * T2;
        if (thr != 256) continue;  // 256 threads x 4 CTAs per SM: the tile's load -> compute chain overlaps across CTAs
        if ((T0 > 1 && T0 / 2 >= od[0]) || (T1 > 1 && T1 / 2 >= od[1]) || (T2 > 8 && T2 / 2 >= od[2])) continue;
        const int T[3] = {T0, T1, T2};
        long long boxv = 1;
        int S[3];
        for (int i = 0; i < 3; ++i) {
          S[i] = (int)ceil((T[i] - 1) * sc[i]) + 3;  // floor difference of the end coordinates <= ceil(span) + 1, plus the far corner
          boxv *= S[i];
        }
        if (boxv * 4 > budget) continue;
        const int CC = (int)std::min<long long>(a.channels, budget / (boxv * 4));
        const double passes = ceil((double)a.channels / CC);
        // staged input voxels per output voxel, a barrier pair per pass, partial tiles at the far edges
        double waste = 1.0;
        for (int i = 0; i < 3; ++i) waste *= (double)((od[i] + T[i] - 1) / T[i] * T[i]) / od[i];
        const double lanes = (double)S[2] / ((S[2] + 31) / 32 * 32);  // lane efficiency of the row loads
        const double cost = ((double)boxv / thr / lanes + 0.25 * passes / a.channels) * waste;
        if (cost < best) {
          best = cost;
          p.T0 = T0, p.T1 = T1, p.T2 = T2, p.S0 = S[0], p.S1 = S[1], p.S2 = S[2], p.CC = CC;
        }
      }
  if (best >= 1e30) return false;
  p.nt0 = (od[0] + p.T0 - 1) / p.T0, p.nt1 = (od[1] + p.T1 - 1) / p.T1, p.nt2 = (od[2] + p.T2 - 1) / p.T2;
  return true;
}

template <bool ARGMAX>
static int launch_trilinear(const TriArgs& a, cudaStream_t st) {
  BrickPlan p;
  const long long ovox = (long long)a.od0 * a.od1 * a.od2;
  if (brick_plan(a, p)) {
    static DeviceOnce attr_once;
    if (attr_once.first()) {
      SGM_CUDA_CHECK(cudaFuncSetAttribute(trilinear_brick_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024));
      SGM_CUDA_CHECK(cudaFuncSetAttribute(trilinear_brick_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024));
    }
    const long long ntiles = (long long)p.nt0 * p.nt1 * p.nt2;
    const size_t smem = (size_t)p.CC * p.S0 * p.S1 * p.S2 * sizeof(float);
    const int grid = (int)std::min<long long>(ntiles, 148LL * 32);
    trilinear_brick_kernel<ARGMAX><<<grid, p.T0 * p.T1 * p.T2, smem, st>>>(a, p);
  } else if (getenv("SGM_RESAMPLE_SEP") && tri_diagonal(a) && !getenv("SGM_NO_RESAMPLE_SEP")) {
    // OPT-IN: on the configs[2] forward Spacing (1 channel) the row kernel measured 3.5 ms against 0.38 ms of the
    // gather kernel -- kept for the equality tests and further work, not on the default path.
    return launch_trilinear_sep<ARGMAX>(a, st);
  } else {
    trilinear_kernel<ARGMAX><<<grid_for(ovox), 256, 0, st>>>(a);
  }
  SGM_CUDA_CHECK(cudaGetLastError());
  return SGM_OK;
}

}  // namespace
}  // namespace sgm

using namespace sgm;

extern "C" int32_t sgm_resample_trilinear(const float* in_dev, const int32_t in_dims[3], int32_t channels,
                                          float* out_dev, const int32_t out_dims[3],
                                          const double xform[12], void* stream) {
  SGM_REQUIRE(in_dev && out_dev && channels > 0, SGM_ERR_INVALID, "resample_trilinear: null argument");
  TriArgs a;
  a.in = in_dev, a.out = out_dev, a.labels = nullptr, a.channels = channels;
  a.id0 = in_dims[0], a.id1 = in_dims[1], a.id2 = in_dims[2];
  a.od0 = out_dims[0], a.od1 = out_dims[1], a.od2 = out_dims[2];
  for (int i = 0; i < 12; ++i) a.m[i] = xform[i];
  return launch_trilinear<false>(a, (cudaStream_t)stream);
}

extern "C" int32_t sgm_resample_trilinear_argmax(const float* in_dev, const int32_t in_dims[3],
                                                 int32_t channels, uint8_t* out_dev,
                                                 const int32_t out_dims[3], const double xform[12],
                                                 void* stream) {
  SGM_REQUIRE(in_dev && out_dev && channels > 0 && channels <= 256, SGM_ERR_INVALID,
              "resample_trilinear_argmax: bad argument");
  TriArgs a;
  a.in = in_dev, a.out = nullptr, a.labels = out_dev, a.channels = channels;
  a.id0 = in_dims[0], a.id1 = in_dims[1], a.id2 = in_dims[2];
  a.od0 = out_dims[0], a.od1 = out_dims[1], a.od2 = out_dims[2];
  for (int i = 0; i < 12; ++i) a.m[i] = xform[i];
  return launch_trilinear<true>(a, (cudaStream_t)stream);
}

extern "C" int32_t sgm_resample_itk(const void* in_dev, int32_t dtype, const int32_t in_dims[3],
                                    void* out_dev, const int32_t out_dims[3],
                                    const double out_index_to_phys[9], const double out_origin[3],
                                    const double in_phys_to_index[9], const double in_origin[3],
                                    int32_t nearest, double default_value, void* stream) {
  SGM_REQUIRE(in_dev && out_dev, SGM_ERR_INVALID, "resample_itk: null argument");
  ItkArgs a;
  a.in = in_dev, a.out = out_dev;
  for (int i = 0; i < 3; ++i) {
    a.in_n[i] = in_dims[i], a.out_n[i] = out_dims[i];
    a.oorg[i] = out_origin[i], a.iorg[i] = in_origin[i];
  }
  for (int i = 0; i < 9; ++i) a.i2p[i] = out_index_to_phys[i], a.p2i[i] = in_phys_to_index[i];
  a.nearest = nearest, a.defval = default_value;
  const long long ovox = (long long)a.out_n[0] * a.out_n[1] * a.out_n[2];
  cudaStream_t st = (cudaStream_t)stream;
  const bool no_vec = getenv("SGM_NO_RESAMPLE_VEC") != nullptr;  // A/B switch (tests compare the two kernels for equality)
  // axis-aligned nearest neighbour (the label path): separable index tables + a pure gather, bit-identical
  // OPT-IN (SGM_RESAMPLE_SEP=1): measured on the configs[2] label map it is SLOWER than the 4-voxels-per-thread kernel
  // (0.27 vs 0.14 ms: the per-call table allocation and kernel cost more than the float64 work they save).
  const bool sep_on = getenv("SGM_RESAMPLE_SEP") != nullptr;
  if (sep_on && nearest && !no_vec && !getenv("SGM_NO_RESAMPLE_SEP") && itk_diagonal(a) && a.out_n[0] <= 12288 &&
      (long long)a.out_n[1] * a.out_n[2] < (1LL << 31)) {
    switch (dtype) {
      case 0: return launch_itk_sep<uint8_t>(a, st);
      case 1: return launch_itk_sep<int16_t>(a, st);
      case 2: return launch_itk_sep<uint16_t>(a, st);
      case 3: return launch_itk_sep<float>(a, st);
      case 4: return launch_itk_sep<int32_t>(a, st);
      default: set_error("resample_itk: unsupported dtype code %d", dtype); return SGM_ERR_UNSUPPORTED;
    }
  }
  const bool vec = !no_vec && a.out_n[0] % 4 == 0 && ((uintptr_t)out_dev % 16) == 0;
  const int g = grid_for(vec ? ovox / 4 : ovox);
  switch (dtype) {
    case 0: if (vec) itk_resample_vec_kernel<uint8_t, 4><<<g, 256, 0, st>>>(a); else itk_resample_kernel<uint8_t><<<g, 256, 0, st>>>(a); break;
    case 1: if (vec) itk_resample_vec_kernel<int16_t, 4><<<g, 256, 0, st>>>(a); else itk_resample_kernel<int16_t><<<g, 256, 0, st>>>(a); break;
    case 2: if (vec) itk_resample_vec_kernel<uint16_t, 4><<<g, 256, 0, st>>>(a); else itk_resample_kernel<uint16_t><<<g, 256, 0, st>>>(a); break;
    case 3: if (vec) itk_resample_vec_kernel<float, 4><<<g, 256, 0, st>>>(a); else itk_resample_kernel<float><<<g, 256, 0, st>>>(a); break;
    case 4: if (vec) itk_resample_vec_kernel<int32_t, 4><<<g, 256, 0, st>>>(a); else itk_resample_kernel<int32_t><<<g, 256, 0, st>>>(a); break;
    default: set_error("resample_itk: unsupported dtype code %d", dtype); return SGM_ERR_UNSUPPORTED;
  }
  SGM_CUDA_CHECK(cudaGetLastError());
  return SGM_OK;
}

extern "C" int32_t sgm_normalize_intensity(const float* in_dev, float* out_dev, int32_t channels,
                                           int64_t voxels, double* scratch_dev, void* stream) {
  SGM_REQUIRE(in_dev && out_dev && scratch_dev && channels > 0 && voxels > 0, SGM_ERR_INVALID,
              "normalize_intensity: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  int nb = grid_for(voxels);
  if (nb > RED_BLOCKS) nb = RED_BLOCKS;
  double* partial = scratch_dev;           // [RED_BLOCKS]
  double* mean_std = scratch_dev + 2048;   // [2]
  for (int c = 0; c < channels; ++c) {
    const float* x = in_dev + (long long)c * voxels;
    sum_kernel<<<nb, 256, 0, st>>>(x, voxels, nullptr, partial);
    reduce_final_kernel<<<1, 32, 0, st>>>(partial, nb, voxels, 0, mean_std);
    sum_kernel<<<nb, 256, 0, st>>>(x, voxels, mean_std, partial);
    reduce_final_kernel<<<1, 32, 0, st>>>(partial, nb, voxels, 1, mean_std + 1);
    normalize_apply_kernel<<<grid_for(voxels), 256, 0, st>>>(x, out_dev + (long long)c * voxels, voxels,
                                                             mean_std);
  }
  SGM_CUDA_CHECK(cudaGetLastError());
  return SGM_OK;
}

extern "C" int32_t sgm_foreground_bbox(const float* in_dev, int32_t channels, const int32_t dims[3],
                                       int32_t* bbox_dev, void* stream) {
  SGM_REQUIRE(in_dev && bbox_dev && channels > 0, SGM_ERR_INVALID, "foreground_bbox: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const long long vox = (long long)dims[0] * dims[1] * dims[2];
  bbox_init_kernel<<<1, 32, 0, st>>>(bbox_dev);
  bbox_kernel<<<grid_for(vox), 256, 0, st>>>(in_dev, channels, dims[0], dims[1], dims[2], bbox_dev);
  bbox_final_kernel<<<1, 32, 0, st>>>(bbox_dev);
  SGM_CUDA_CHECK(cudaGetLastError());
  return SGM_OK;
}
