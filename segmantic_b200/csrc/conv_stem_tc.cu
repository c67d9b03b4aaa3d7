// Tensor-core stem: the first down block of the UNet (unit0 conv k3 s2 + the residual-branch conv k3 s2, both
// reading the 1-2 channel network input) as ONE tcgen05 GEMM with an explicit im2col tile.
//
// With a single input channel there is no channel dimension to contract over, so the implicit-GEMM trick of the
// other layers (A operand = shifted views of a halo brick) does not apply: K must hold the 27 taps themselves.
// A CTA stages the fp32 halo brick of a 4x4x8 output tile (windows are read straight from the planar volume,
// zero outside the WINDOW as MONAI convolves every window in isolation), every thread builds the im2col row of
// its output voxel in bf16 (K = 27 * Cin padded to a multiple of 16), one elected thread issues K/16 MMAs
// (M = 128 voxels, N = 32 = 16 unit0 + 16 residual channels), and the four warps turn the accumulator into the two
// bf16 CG8 tensors.  On CUDA cores this layer was FFMA-bound at 11 us per window (8 % of the step).
// Rounding: the network input is rounded to bf16 here (oracle/bf16_emulation.py rounds at the same point).
#include "common.cuh"
#include "tc_ptx.cuh"

#include <stdlib.h>
#include <string.h>

#include <vector>

namespace sgm {

namespace {
using namespace tcptx;

constexpr int T0 = 4, T1 = 4, T2 = 8;                      // output tile: 128 voxels = the 128 GEMM rows
constexpr int B0 = 2 * T0 + 1, B1 = 2 * T1 + 1, B2 = 2 * T2 + 1;  // halo brick of the stride-2 k3 conv
constexpr int BRICK = B0 * B1 * B2;
constexpr int kMaxK = 64;                                  // 27 * Cin <= 64: one or two input channels

struct StemTcArgs {
  const float* vol;
  long long vol_cstride;
  int vd1, vd2;
  const int* win_origin;  // device int[n][3]
  int cin, n;
  int id[3], od[3];
  int nt[3], ntiles;      // tiles per window
  int KP;                 // padded K
  const __nv_bfloat16* w; // [KP/8][32][8]
  const float* bias;      // [32]: unit0 (16, zero padded) then residual branch (16)
  __nv_bfloat16* outA;
  __nv_bfloat16* outB;
  int cgA, cgB, actA;
  float alphaA;
  int* error_flag;
};

template <int CIN>
__global__ void __launch_bounds__(128, 6) stem_tc_kernel(const StemTcArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5;
  constexpr int KP = (27 * CIN + 15) / 16 * 16;  // padded K
  constexpr int nkc = KP / 8;                    // 16-byte K chunks per row
  uint8_t* a_tile = smem;                                   // [nkc][128 rows][16 B], K-major SWIZZLE_NONE
  uint8_t* b_tile = a_tile + (size_t)nkc * 2048;            // [nkc][32 rows][16 B]
  float* brick = reinterpret_cast<float*>(b_tile + (size_t)nkc * 512);  // [cin][B0][B1][B2] fp32
  uint64_t* bar = reinterpret_cast<uint64_t*>(brick + (size_t)CIN * BRICK + (CIN * BRICK & 1));
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  float* bias_s = reinterpret_cast<float*>(bar + 2);
  const uint32_t bar_u = smem_u32(bar);

  if (tid == 0) {
    mbar_init(bar_u, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(32u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = tid; i < nkc * 32; i += 128)  // weights: resident for the whole kernel
    reinterpret_cast<uint4*>(b_tile)[i] = __ldg(reinterpret_cast<const uint4*>(a.w) + i);
  if (tid < 32) bias_s[tid] = __ldg(a.bias + tid);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  // this thread's output voxel inside the tile, and the brick offset of its tap (0, 0, 0)
  const int l0 = tid >> 5, l1 = (tid >> 3) & 3, l2 = tid & 7;
  const int boff = (2 * l0 * B1 + 2 * l1) * B2 + 2 * l2;
  const long long ovox = (long long)a.od[0] * a.od[1] * a.od[2];
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
  const long long total = (long long)a.n * a.ntiles;
  uint32_t phase = 0;
  // The halo brick of tile i + 1 is fetched into registers while tile i is converted, multiplied and stored: the
  // kernel was bound by the latency of these loads (long-scoreboard stalls, 37 % active warps).
  constexpr int NLD = (CIN * BRICK + 127) / 128;
  float ld[NLD];
  auto fetch = [&](long long tile) {
    const int n = (int)(tile / a.ntiles);
    int r = (int)(tile - (long long)n * a.ntiles);
    const int b2 = r % a.nt[2];
    r /= a.nt[2];
    const int b1 = r % a.nt[1], b0 = r / a.nt[1];
    const int w0 = a.win_origin[n * 3 + 0], w1 = a.win_origin[n * 3 + 1], w2 = a.win_origin[n * 3 + 2];
    // zero outside the window (input coordinate = 2 * output - 1 + brick index)
#pragma unroll
    for (int it = 0; it < NLD; ++it) {
      const int i = tid + it * 128;
      const int ci = i / BRICK;
      int e = i - ci * BRICK;
      const int e2 = e % B2;
      e /= B2;
      const int e1 = e % B1, e0 = e / B1;
      const int i0 = 2 * b0 * T0 - 1 + e0, i1 = 2 * b1 * T1 - 1 + e1, i2 = 2 * b2 * T2 - 1 + e2;
      ld[it] = 0.f;
      if (i < CIN * BRICK && i0 >= 0 && i0 < a.id[0] && i1 >= 0 && i1 < a.id[1] && i2 >= 0 && i2 < a.id[2])
        ld[it] = __ldg(a.vol + ci * a.vol_cstride + ((long long)(w0 + i0) * a.vd1 + (w1 + i1)) * a.vd2 + (w2 + i2));
    }
  };
  if (blockIdx.x < total) fetch(blockIdx.x);
  for (long long tile = blockIdx.x; tile < total; tile += gridDim.x, phase ^= 1u) {
    const int n = (int)(tile / a.ntiles);
    int r = (int)(tile - (long long)n * a.ntiles);
    const int b2 = r % a.nt[2];
    r /= a.nt[2];
    const int b1 = r % a.nt[1], b0 = r / a.nt[1];
    const int o0b = b0 * T0, o1b = b1 * T1, o2b = b2 * T2;
#pragma unroll
    for (int it = 0; it < NLD; ++it)
      if (tid + it * 128 < CIN * BRICK) brick[tid + it * 128] = ld[it];
    if (tile + gridDim.x < total) fetch(tile + gridDim.x);
    __syncthreads();
    // ---- im2col row of this thread's voxel: K index = ci * 27 + tap, bf16, 8 values per 16-byte chunk
#pragma unroll
    for (int kc = 0; kc < nkc; ++kc) {
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int kk = kc * 8 + j;
        float x = 0.f;
        if (kk < CIN * 27) {
          const int ci = kk / 27, tap = kk - ci * 27;
          const int k0 = tap / 9, k1 = (tap / 3) % 3, k2 = tap % 3;
          x = brick[ci * BRICK + boff + (k0 * B1 + k1) * B2 + k2];
        }
        v[j] = x;
      }
      *reinterpret_cast<uint4*>(a_tile + (size_t)kc * 2048 + (size_t)tid * 16) = pack8(v);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> tensor-core reads
    tc_fence_before();
    __syncthreads();
    if (warp == 0 && elect_one()) {
      tc_fence_after();
      const uint32_t a16 = smem_u32(a_tile) >> 4, b16 = smem_u32(b_tile) >> 4;
#pragma unroll
      for (int ks = 0; ks < KP / 16; ++ks)  // K = 16 per MMA: chunks 2 ks and 2 ks + 1
        tc_mma(tmem, make_desc(a16 + ks * 256, 128, 8), make_desc(b16 + ks * 64, 32, 8), idesc, ks > 0 ? 1u : 0u);
      tc_commit(bar_u);
    }
    __syncwarp();
    if (!mbar_wait(bar_u, phase, a.error_flag, 41)) break;
    tc_fence_after();
    uint32_t raw[32];
    tc_ld16(tmem + ((uint32_t)(warp * 32) << 16), raw);
    tc_ld16(tmem + ((uint32_t)(warp * 32) << 16) + 16, raw + 16);
    tc_fence_before();
    const int o0 = o0b + l0, o1 = o1b + l1, o2 = o2b + l2;
    if (o0 < a.od[0] && o1 < a.od[1] && o2 < a.od[2]) {
      const long long opos = ((long long)o0 * a.od[1] + o1) * a.od[2] + o2;
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        float va[8], vb[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float x = __uint_as_float(raw[8 * g + c]) + bias_s[8 * g + c];
          if (a.actA) x = prelu(x, a.alphaA);
          va[c] = x;
          vb[c] = __uint_as_float(raw[16 + 8 * g + c]) + bias_s[16 + 8 * g + c];
        }
        if (g < a.cgA) *reinterpret_cast<uint4*>(a.outA + (((long long)n * a.cgA + g) * ovox + opos) * 8) = pack8(va);
        if (g < a.cgB) *reinterpret_cast<uint4*>(a.outB + (((long long)n * a.cgB + g) * ovox + opos) * 8) = pack8(vb);
      }
    }
    __syncthreads();  // the accumulator, the brick and the im2col tile are free again
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32u) : "memory");
}

inline uint16_t f2bf(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7f800000u) != 0x7f800000u) u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}

}  // namespace

// Packs unit0 + residual-branch conv of a 3-D stride-2 k3 stem for the tensor-core kernel; leaves *w_dev null when
// the block is not eligible (the CUDA-core stem_kernel handles it).
int stem_tc_pack(const sgm_conv_desc& u0, const sgm_conv_desc& rs, int spatial_dims, void** w_dev, float** b_dev, int* kp) {
  *w_dev = nullptr, *b_dev = nullptr, *kp = 0;
  if (getenv("SGM_NO_STEM_TC")) return SGM_OK;
  if (spatial_dims != 3 || u0.kind != SGM_KIND_CONV || rs.kind != SGM_KIND_CONV) return SGM_OK;
  if (u0.kernel != 3 || u0.stride != 2 || rs.kernel != 3 || rs.stride != 2) return SGM_OK;
  if (u0.cin > 2 || u0.cin * 27 > kMaxK || u0.cout > 16 || rs.cout > 16 || rs.cin != u0.cin) return SGM_OK;
  const int K = u0.cin * 27, KP = (K + 15) / 16 * 16;
  std::vector<uint16_t> w((size_t)KP * 32, 0);
  std::vector<float> b(32, 0.f);
  for (int f = 0; f < 32; ++f) {
    const sgm_conv_desc& src = f < 16 ? u0 : rs;
    const int co = f < 16 ? f : f - 16;
    if (co >= src.cout) continue;
    b[f] = src.bias[co];
    for (int ci = 0; ci < u0.cin; ++ci)
      for (int tap = 0; tap < 27; ++tap) {
        const int kk = ci * 27 + tap;
        w[((size_t)(kk >> 3) * 32 + f) * 8 + (kk & 7)] = f2bf(src.weight[((size_t)co * src.cin + ci) * 27 + tap]);
      }
  }
  if (cudaMalloc(w_dev, w.size() * 2) != cudaSuccess || cudaMalloc(b_dev, b.size() * 4) != cudaSuccess) {
    set_error("stem_tc_pack: cudaMalloc failed");
    return SGM_ERR_CUDA;
  }
  SGM_CUDA_CHECK(cudaMemcpy(*w_dev, w.data(), w.size() * 2, cudaMemcpyHostToDevice));
  SGM_CUDA_CHECK(cudaMemcpy(*b_dev, b.data(), b.size() * 4, cudaMemcpyHostToDevice));
  *kp = KP;
  return SGM_OK;
}

int launch_stem_tc(const ConvArgs& a, const void* w_dev, const float* b_dev, int kp, int cgA, int cgB, void* outA,
                   void* outB, int* error_flag_dev, cudaStream_t st) {
  StemTcArgs s;
  memset(&s, 0, sizeof(s));
  s.vol = reinterpret_cast<const float*>(a.in0), s.vol_cstride = a.vol_cstride, s.vd1 = a.vd1, s.vd2 = a.vd2;
  s.win_origin = a.win_origin, s.cin = a.cin_real, s.n = a.n;
  for (int i = 0; i < 3; ++i) s.id[i] = a.id[i], s.od[i] = a.od[i];
  s.nt[0] = ceil_div(a.od[0], T0), s.nt[1] = ceil_div(a.od[1], T1), s.nt[2] = ceil_div(a.od[2], T2);
  s.ntiles = s.nt[0] * s.nt[1] * s.nt[2];
  s.KP = kp;
  s.w = reinterpret_cast<const __nv_bfloat16*>(w_dev), s.bias = b_dev;
  s.outA = reinterpret_cast<__nv_bfloat16*>(outA), s.outB = reinterpret_cast<__nv_bfloat16*>(outB);
  s.cgA = cgA, s.cgB = cgB, s.actA = a.act, s.alphaA = a.alpha;
  s.error_flag = error_flag_dev;
  const int nkc = kp / 8;
  const size_t smem = (size_t)nkc * 2048 + (size_t)nkc * 512 + ((size_t)s.cin * BRICK + 2) * 4 + 16 + 128 + 16;
  const long long total = (long long)s.n * s.ntiles;
  const int grid = (int)std::min<long long>(total, 148LL * 6);  // persistent: one wave of 6 CTAs per SM
  if (s.cin == 1) stem_tc_kernel<1><<<grid, 128, smem, st>>>(s);
  else stem_tc_kernel<2><<<grid, 128, smem, st>>>(s);
  SGM_CUDA_CHECK(cudaGetLastError());
  return SGM_OK;
}

}  // namespace sgm
