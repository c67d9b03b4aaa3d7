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
    double cs[3], ce[3], c[3];
    itk_cindex(a, 0.0, (double)oy, (double)oz, cs);
    itk_cindex(a, (double)a.out_n[0], (double)oy, (double)oz, ce);
    const double alpha = (double)ox / (double)a.out_n[0];
    bool inside = true;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      c[r] = __dadd_rn(cs[r], __dmul_rn(alpha, __dadd_rn(ce[r], -cs[r])));
      inside = inside && (c[r] >= -0.5) && (c[r] < (double)a.in_n[r] - 0.5);
    }
    T res;
    if (!inside) {
      res = itk_cast<T>(a.defval);
    } else if (a.nearest) {
      int idx[3];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        int i = (int)floor(c[r] + 0.5);
        idx[r] = min(max(i, 0), a.in_n[r] - 1);
      }
      res = in[((long long)idx[2] * a.in_n[1] + idx[1]) * a.in_n[0] + idx[0]];
    } else {
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
      res = itk_cast<T>(__dadd_rn(vy0, __dmul_rn(fr[2], __dadd_rn(vy1, -vy0))));
    }
    out[v] = res;
  }
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
  const long long ovox = (long long)a.od0 * a.od1 * a.od2;
  trilinear_kernel<false><<<grid_for(ovox), 256, 0, (cudaStream_t)stream>>>(a);
  SGM_CUDA_CHECK(cudaGetLastError());
  return SGM_OK;
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
  const long long ovox = (long long)a.od0 * a.od1 * a.od2;
  trilinear_kernel<true><<<grid_for(ovox), 256, 0, (cudaStream_t)stream>>>(a);
  SGM_CUDA_CHECK(cudaGetLastError());
  return SGM_OK;
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
  const int g = grid_for(ovox);
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case 0: itk_resample_kernel<uint8_t><<<g, 256, 0, st>>>(a); break;
    case 1: itk_resample_kernel<int16_t><<<g, 256, 0, st>>>(a); break;
    case 2: itk_resample_kernel<uint16_t><<<g, 256, 0, st>>>(a); break;
    case 3: itk_resample_kernel<float><<<g, 256, 0, st>>>(a); break;
    case 4: itk_resample_kernel<int32_t><<<g, 256, 0, st>>>(a); break;
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
