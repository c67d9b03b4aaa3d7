// Channel-streamed, PERSISTENT tcgen05 brick kernel for the coarse half of the U (the layers at 12^3 and 6^3 of a
// 96^3 window: 64..384 input channels, 32..256 output channels; stride 1, stride 2 and transposed stride 2).
//
// Why.  Those layers are GEMM-shaped (K = 27 * Cin = 1728..10368) but tiny in M (216 or 1728 voxels per window), and
// the one-CTA-per-brick kernel (conv_tc.cu) ran them at 2-19 % of the dense bf16 peak: TMEM allocation, brick load,
// a short MMA burst and the epilogue ran in sequence inside a CTA, every CTA re-streamed the whole filter bank for 1-3
// row tiles, and 64-channel bricks had to be so small to fit shared memory that ~80 % of the GEMM rows were halo.
//
// How.  One CTA per SM lives for the whole layer and walks a list of UNITS = (group of G windows, brick, block of N
// output channels).  The K dimension is streamed in chunks of 16 input channels: per chunk the A producer loads the
// chunk's halo brick of every window of the unit (TMA boxes with hardware zero fill = the conv's zero padding at the
// window border; stride 2: eight parity slabs by strided boxes; transposed: the input brick) into a ring, the W
// producer streams the chunk's filter taps through a second ring, and the issuers run  taps x row tiles  MMAs
// (M = 128 padded-linear brick rows, N = 64/128 columns, K = 16) into accumulators that stay in TMEM for the whole K
// loop.  All row tiles of all G windows share every weight stage, so the filter bank is read G * tiles times less
// often than before; TMEM is double-buffered by unit (two halves of 256 columns) so the epilogue of a unit (TMEM ->
// bias / PReLU / residual -> bf16 CG8 stores) overlaps the MMAs of the next; barriers and TMEM are set up once.
// Four issuer threads split the row tiles of a unit (tile % 4), so every accumulator receives its MMAs from ONE
// thread in a fixed order -- results are bit-identical from run to run and for every window batch -- while the ~6 clk
// per dependent instruction of a single issuing thread no longer gates the tensor pipe (measured: 49 clk per N = 64 MMA
// against 175 in the one-CTA-per-brick kernel; profiles/r02_cs_notes.md).
//
// MMA cost (tests/ubench_mma.cu): max(N / 2, 32 + N / 4) clk for M = 128, K = 16 -- shared-memory operand reads below
// N = 128, the tensor pipe above: N = 64 runs at 67 % and N = 128 at 100 % of the tensor rate.
#include "conv_tc.cuh"
#include "tc_ptx.cuh"

#include <algorithm>
#include <stdlib.h>
#include <string.h>

namespace sgm {
namespace tc {

namespace {
using namespace tcptx;

constexpr int kEpiWarps = 8;                      // two groups x the four TMEM lane quarters
constexpr int kIssuers = 4;
constexpr int kThreads = (kEpiWarps + 2 + kIssuers) * 32;   // + A producer + W producer
constexpr int kSmemMax = 227 * 1024;
constexpr int kMaxTaps = 32;
constexpr int kMaxStages = 6;
constexpr int kMaxTilesPerIssuer = 2;             // 512 TMEM columns / 64 columns per tile / 4 issuers

struct CsArgs {
  int mode;
  int cg0, cg1, nkc;            // channel groups of the two inputs; 16-channel chunks
  int id[3], od[3], rd[3];      // input / output / row-space extents
  int t[3], H[3], lo[3], nt[3], par[3];
  int P, nslab, row_first, ntiles, G;
  int N, NB, ncls, ntap, ncoblk;
  int nbuf;
  int astages, a_stage_units, a_g_units, box_bytes;
  int wstages, w_stage_bytes, Gw, ngw;
  int nwin, nbricks, nunits;
  const __nv_bfloat16* w;
  const float* bias;
  __nv_bfloat16* outA;
  int cgA;
  __nv_bfloat16* outB;
  int cgB;
  int segA_cg, actA;
  float alphaA;
  const __nv_bfloat16* res;     // global CG8 residual added to segment A after the activation, or null
  int* error_flag;
  long long* trace;
  uint32_t tap_a[kMaxTaps];     // A offset of the tap inside a window's chunk brick (16-byte units, row_first included)
  uint32_t tap_c[kMaxTaps];     // accumulator column offset of the tap's class | first-of-class << 31
};

struct Unit {
  int coblk, w0, gn, org[3];
};

__device__ __forceinline__ Unit decode_unit(const CsArgs& a, int u) {
  Unit r;
  r.coblk = u % a.ncoblk;
  int q = u / a.ncoblk;
  const int brick = q % a.nbricks;
  const int wg = q / a.nbricks;
  r.w0 = wg * a.G;
  r.gn = min(a.G, a.nwin - r.w0);
  const int b2 = brick % a.nt[2];
  const int b01 = brick / a.nt[2];
  r.org[0] = (b01 / a.nt[1]) * a.t[0], r.org[1] = (b01 % a.nt[1]) * a.t[1], r.org[2] = b2 * a.t[2];
  return r;
}

struct IssueCtx {
  uint32_t a_base16, w_base16, w_stage16, a_lbo, b_lbo, desc_hi, idesc, col_base, bar0;
};

// The K loop of one unit for an issuer that owns NMY row tiles (compile-time: no predicates around the MMAs).
// Per 16-channel chunk: wait for the A stage; per weight stage: wait, then  taps x tiles  MMAs; commits release the
// stages.  An issuer without tiles (NMY = 0) still waits and commits so that every barrier sees all its arrivals.
template <int NMY>
__device__ __forceinline__ void issue_unit(const CsArgs& a, const IssueCtx& ic, const uint32_t* tile_a, const uint32_t* tile_c,
                                           int& sa, uint32_t& pa, int& sw, uint32_t& pw) {
  const uint32_t NB2 = (uint32_t)(a.NB * 2);
  uint32_t col[NMY > 0 ? NMY : 1];
#pragma unroll
  for (int q = 0; q < NMY; ++q) col[q] = ic.col_base + tile_c[q];
  for (int kc = 0; kc < a.nkc; ++kc) {
    mbar_wait_or_trap(ic.bar0 + 8u * sa, pa, a.error_flag, 54);  // AFULL(sa)
    uint32_t a_tile[NMY > 0 ? NMY : 1];
#pragma unroll
    for (int q = 0; q < NMY; ++q) a_tile[q] = ((ic.a_base16 + (uint32_t)(sa * a.a_stage_units)) | ic.a_lbo) + tile_a[q];
    int tap = 0;
    for (int gi = 0; gi < a.ngw; ++gi) {
      mbar_wait_or_trap(ic.bar0 + 8u * (2 * kMaxStages + sw), pw, a.error_flag, 55);  // WFULL(sw)
      tc_fence_after();
      uint32_t b_lo = (ic.w_base16 + (uint32_t)sw * ic.w_stage16) | ic.b_lbo;
      const int nb = min(a.Gw, a.ntap - gi * a.Gw);
#pragma unroll 3
      for (int j = 0; j < nb; ++j, ++tap, b_lo += NB2) {
        const uint32_t ta = a.tap_a[tap];
        const uint32_t acc = (kc | tap) ? 1u : 0u;  // the very first MMA of a unit overwrites the accumulator
        const uint64_t bdesc = ((uint64_t)ic.desc_hi << 32) | b_lo;
#pragma unroll
        for (int q = 0; q < NMY; ++q)
          tc_mma(col[q], ((uint64_t)ic.desc_hi << 32) | (a_tile[q] + ta), bdesc, ic.idesc, acc);
      }
      tc_commit(ic.bar0 + 8u * (3 * kMaxStages + sw));  // WEMPTY(sw)
      if (++sw == a.wstages) sw = 0, pw ^= 1u;
    }
    tc_commit(ic.bar0 + 8u * (kMaxStages + sa));  // AEMPTY(sa)
    if (++sa == a.astages) sa = 0, pa ^= 1u;
  }
}

__global__ void __launch_bounds__(kThreads, 1)
cs_conv_kernel(const __grid_constant__ CsArgs a, const __grid_constant__ CUtensorMap tmap0,
               const __grid_constant__ CUtensorMap tmap1) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t a_stage_bytes = (uint32_t)a.a_stage_units * 16u;
  uint8_t* a_smem = smem;
  uint8_t* w_smem = smem + (size_t)a.astages * a_stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(w_smem + (size_t)a.wstages * a.w_stage_bytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4 * kMaxStages + 4);
  const uint32_t bar0 = smem_u32(bars);
  auto AFULL = [&](int s) { return bar0 + 8u * s; };
  auto AEMPTY = [&](int s) { return bar0 + 8u * (kMaxStages + s); };
  auto WFULL = [&](int s) { return bar0 + 8u * (2 * kMaxStages + s); };
  auto WEMPTY = [&](int s) { return bar0 + 8u * (3 * kMaxStages + s); };
  auto TFULL = [&](int b) { return bar0 + 8u * (4 * kMaxStages + b); };
  auto TEMPTY = [&](int b) { return bar0 + 8u * (4 * kMaxStages + 2 + b); };
  const bool tr = a.trace != nullptr && blockIdx.x == 0;

  if (tid == 0) {
    for (int s = 0; s < kMaxStages; ++s) {
      mbar_init(AFULL(s), 1);
      mbar_init(AEMPTY(s), kIssuers);
      mbar_init(WFULL(s), 1);
      mbar_init(WEMPTY(s), kIssuers);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(TFULL(b), kIssuers);
      mbar_init(TEMPTY(b), kEpiWarps * 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (tr) a.trace[0] = clock64();
  }
  if (warp == kEpiWarps) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int tiles_g = a.ntiles;                  // row tiles per window of the unit
  const int cols_tile = a.ncls * a.N;

  if (warp == kEpiWarps) {
    // ===================== A producer: per unit and 16-channel chunk, the chunk's brick of every window =====================
    if (lane == 0) {
      const uint32_t a_base = smem_u32(a_smem);
      int s = 0;
      uint32_t ph = 0;     // parity of the stage's CURRENT fill
      bool wrapped = false;
      for (int u = blockIdx.x; u < a.nunits; u += gridDim.x) {
        const Unit un = decode_unit(a, u);
        int c0, c1, c2;  // box origin along d2, d1, d0
        if (a.mode == MODE_S2) {
          c0 = un.org[2] * 2 - 2, c1 = un.org[1] * 2 - 2, c2 = un.org[0] * 2 - 2;
        } else {
          c0 = un.org[2] - a.lo[2], c1 = un.org[1] - a.lo[1], c2 = un.org[0] - a.lo[0];
        }
        for (int kc = 0; kc < a.nkc; ++kc) {
          if (wrapped) mbar_wait_or_trap(AEMPTY(s), ph ^ 1u, a.error_flag, 51);  // the previous fill was consumed
          mbar_expect_tx(AFULL(s), (uint32_t)(un.gn * a.nslab * 2 * a.box_bytes));
          const uint32_t dst_s = a_base + (uint32_t)s * a_stage_bytes;
          for (int g = 0; g < un.gn; ++g) {
            const int n = un.w0 + g;
            for (int slab = 0; slab < a.nslab; ++slab) {
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const int cg = kc * 2 + h;
                const bool first = cg < a.cg0;
                const CUtensorMap* tm = first ? &tmap0 : &tmap1;
                const int gidx = first ? n * a.cg0 + cg : n * a.cg1 + (cg - a.cg0);
                const uint32_t dst = dst_s + (uint32_t)(g * a.a_g_units + (slab * 2 + h) * a.P) * 16u;
                if (a.mode == MODE_S2) {
                  const int r2 = slab & 1, r1 = (slab >> 1) & 1, r0 = (slab >> 2) & 1;
                  tma_load_5d(dst, tm, 0, c0 + r2, c1 + r1, c2 + r0, gidx, AFULL(s));
                } else {
                  tma_load_4d(dst, tm, c0 * 8, c1, c2, gidx, AFULL(s));
                }
              }
            }
          }
          if (++s == a.astages) s = 0, ph ^= 1u, wrapped = true;
        }
      }
    }
  } else if (warp == kEpiWarps + 1) {
    // ===================== W producer: the chunk's filter taps, Gw tap blocks per stage =====================
    if (lane == 0) {
      const uint32_t w_base = smem_u32(w_smem);
      int s = 0;
      uint32_t ph = 0;
      bool wrapped = false;
      for (int u = blockIdx.x; u < a.nunits; u += gridDim.x) {
        const int coblk = u % a.ncoblk;
        const __nv_bfloat16* wsrc = a.w + (size_t)coblk * a.nkc * a.ntap * a.NB * 16;
        for (int kc = 0; kc < a.nkc; ++kc) {
          for (int gi = 0; gi < a.ngw; ++gi) {
            if (wrapped) mbar_wait_or_trap(WEMPTY(s), ph ^ 1u, a.error_flag, 52);
            const int nb = min(a.Gw, a.ntap - gi * a.Gw);
            const uint32_t bytes = (uint32_t)nb * a.NB * 32u;
            mbar_expect_tx(WFULL(s), bytes);
            bulk_g2s(w_base + (uint32_t)s * a.w_stage_bytes, wsrc + (size_t)(kc * a.ntap + gi * a.Gw) * a.NB * 16, bytes,
                     WFULL(s));
            if (++s == a.wstages) s = 0, ph ^= 1u, wrapped = true;
          }
        }
      }
    }
  } else if (warp > kEpiWarps + 1) {
    // ===================== MMA issuers: issuer iw owns the unit's row tiles tt with tt % kIssuers == iw =====================
    if (elect_one()) {
      const int iw = warp - kEpiWarps - 2;
      const int NB = a.NB;
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NB >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t a_base16 = smem_u32(a_smem) >> 4;
      const uint32_t w_base16 = smem_u32(w_smem) >> 4;
      const uint32_t desc_hi = 8u | (1u << 14);                 // SBO = 128 B, descriptor version 1
      const uint32_t a_lbo = ((uint32_t)a.P & 0x3FFFu) << 16;   // second channel group of the chunk
      const uint32_t b_lbo = ((uint32_t)NB & 0x3FFFu) << 16;    // second 8 input channels of the tap block
      const uint32_t w_stage16 = (uint32_t)a.w_stage_bytes >> 4;
      int sa = 0, sw = 0, uit = 0;
      uint32_t pa = 0, pw = 0;  // parities of the ring stages' current fills
      for (int u = blockIdx.x; u < a.nunits; u += gridDim.x, ++uit) {
        const int gn = min(a.G, a.nwin - (u / a.ncoblk / a.nbricks) * a.G);
        // this issuer's tiles: A offset inside the stage and accumulator column, kept in registers
        uint32_t tile_a[kMaxTilesPerIssuer], tile_c[kMaxTilesPerIssuer];
        int nmy = 0;
        {
          int tt = 0;
          for (int g = 0; g < gn; ++g)
            for (int t = 0; t < tiles_g; ++t, ++tt)
              if ((tt % kIssuers) == iw) {
#pragma unroll
                for (int q = 0; q < kMaxTilesPerIssuer; ++q)
                  if (q == nmy) tile_a[q] = (uint32_t)(g * a.a_g_units + t * 128), tile_c[q] = (uint32_t)(tt * cols_tile);
                ++nmy;
              }
        }
        const int buf = a.nbuf == 2 ? (uit & 1) : 0;
        const int use = a.nbuf == 2 ? (uit >> 1) : uit;
        if (use > 0) mbar_wait_or_trap(TEMPTY(buf), (uint32_t)(use - 1) & 1u, a.error_flag, 53);
        tc_fence_after();
        const uint32_t col_base = tmem_base + (uint32_t)(buf * 256);
        IssueCtx ic{a_base16, w_base16, w_stage16, a_lbo, b_lbo, desc_hi, idesc, col_base, bar0};
        if (nmy >= 2) issue_unit<2>(a, ic, tile_a, tile_c, sa, pa, sw, pw);
        else if (nmy == 1) issue_unit<1>(a, ic, tile_a, tile_c, sa, pa, sw, pw);
        else issue_unit<0>(a, ic, tile_a, tile_c, sa, pa, sw, pw);
        tc_commit(TFULL(buf));
        if (tr && iw == 0 && uit < 6) a.trace[1 + uit] = clock64();
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue: warp = (group eg, lane quarter q4); group eg owns tiles tt % 2 == eg =====================
    const int q4 = warp & 3, eg = warp >> 2;
    const long long ovox = (long long)a.od[0] * a.od[1] * a.od[2];
    const int N = a.N, npiece = N / 16;
    int uit = 0;
    bool ok = true;
    for (int u = blockIdx.x; u < a.nunits && ok; u += gridDim.x, ++uit) {
      const Unit un = decode_unit(a, u);
      const int buf = a.nbuf == 2 ? (uit & 1) : 0;
      const int use = a.nbuf == 2 ? (uit >> 1) : uit;
      ok = mbar_wait(TFULL(buf), (uint32_t)use & 1u, a.error_flag, 56);
      if (!ok) break;
      tc_fence_after();
      if (tr && warp == 0 && lane == 0 && uit < 6) a.trace[8 + uit] = clock64();
      const int ntt = un.gn * tiles_g;
      for (int tt = eg; tt < ntt; tt += kEpiWarps / 4) {
        const int g = tt / tiles_g, t = tt - g * tiles_g;
        const int n = un.w0 + g;
        const int p = a.row_first + t * 128 + q4 * 32 + lane;
        const int h2 = p % a.H[2];
        const int h01 = p / a.H[2];
        const int h1 = h01 % a.H[1], h0 = h01 / a.H[1];
        const int r0 = un.org[0] + h0 - a.lo[0], r1 = un.org[1] + h1 - a.lo[1], r2 = un.org[2] + h2 - a.lo[2];
        const bool valid = h0 >= a.lo[0] && h0 < a.lo[0] + a.t[0] && h1 >= a.lo[1] && h1 < a.lo[1] + a.t[1] &&
                           h2 >= a.lo[2] && h2 < a.lo[2] + a.t[2] && r0 < a.rd[0] && r1 < a.rd[1] && r2 < a.rd[2];
        for (int cls = 0; cls < a.ncls; ++cls) {
          int o0 = r0, o1 = r1, o2 = r2;
          if (a.mode == MODE_T2) o2 = 2 * r2 + (cls & 1), o1 = 2 * r1 + ((cls >> 1) & 1), o0 = 2 * r0 + ((cls >> 2) & 1);
          const long long opos = ((long long)o0 * a.od[1] + o1) * a.od[2] + o2;
          const uint32_t tcol0 = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(buf * 256 + tt * cols_tile + cls * N);
          uint4 nres0 = make_uint4(0, 0, 0, 0), nres1 = nres0;
          auto res_fetch = [&](int piece) {
            const int gq = (un.coblk * N + piece * 16) >> 3;
            if (valid && gq < a.segA_cg) {
              nres0 = __ldg(reinterpret_cast<const uint4*>(a.res + (((long long)n * a.cgA + gq) * ovox + opos) * 8));
              nres1 = __ldg(reinterpret_cast<const uint4*>(a.res + (((long long)n * a.cgA + gq + 1) * ovox + opos) * 8));
            }
          };
          if (a.res) res_fetch(0);
          for (int piece = 0; piece < npiece; ++piece) {
            uint32_t raw[16];
            tc_ld16(tcol0 + piece * 16, raw);  // warp-collective: every lane executes it
            const uint4 cres0 = nres0, cres1 = nres1;
            if (a.res && piece + 1 < npiece) res_fetch(piece + 1);
            if (!valid) continue;
            const int cbase = un.coblk * N + piece * 16;  // fused output channel of raw[0]
            const int gcg = cbase >> 3;
            const bool segA = gcg < a.segA_cg;
            float v[16];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.bias + cbase) + k);
              v[4 * k + 0] = __uint_as_float(raw[4 * k + 0]) + b4.x;
              v[4 * k + 1] = __uint_as_float(raw[4 * k + 1]) + b4.y;
              v[4 * k + 2] = __uint_as_float(raw[4 * k + 2]) + b4.z;
              v[4 * k + 3] = __uint_as_float(raw[4 * k + 3]) + b4.w;
            }
            if (segA) {
              if (a.actA) {
#pragma unroll
                for (int c = 0; c < 16; ++c) v[c] = prelu(v[c], a.alphaA);
              }
              if (a.res) {
                float r[8];
                unpack8(cres0, r);
#pragma unroll
                for (int c = 0; c < 8; ++c) v[c] += r[c];
                unpack8(cres1, r);
#pragma unroll
                for (int c = 0; c < 8; ++c) v[8 + c] += r[c];
              }
#pragma unroll
              for (int h = 0; h < 2; ++h)
                if (gcg + h < a.cgA)
                  *reinterpret_cast<uint4*>(a.outA + (((long long)n * a.cgA + gcg + h) * ovox + opos) * 8) = pack8(v + h * 8);
            } else {
              const int bcg = gcg - a.segA_cg;
#pragma unroll
              for (int h = 0; h < 2; ++h)
                if (bcg + h < a.cgB)
                  *reinterpret_cast<uint4*>(a.outB + (((long long)n * a.cgB + bcg + h) * ovox + opos) * 8) = pack8(v + h * 8);
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(TEMPTY(buf));
      if (tr && warp == 0 && lane == 0 && uit < 6) a.trace[14 + uit] = clock64();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kEpiWarps) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
  if (tr && tid == 0) a.trace[20] = clock64();
}

// ------------------------------------------------------------------------------------------- host
inline uint16_t f2bf(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7f800000u) != 0x7f800000u) u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}

inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    if (const char* env = getenv("SGM_RESERVE_SMS")) n = std::max(8, n - atoi(env));
  }
  return n;
}

struct CsTap {
  int shift[3], slab, first;
};

struct CsPack {
  int N, NB, ncoblk, ntap, nkc;
  std::vector<CsTap> taps;
};

struct CsPlan {
  int key[8];
  CsArgs args;
  int smem_bytes, grid;
};

// MMA cost model of the header (clk per M = 128, K = 16 MMA)
inline double mma_clk(int nb) { return std::max(nb / 2.0, 32.0 + nb / 4.0); }

}  // namespace

struct CsState {
  CsPack pack;
  std::vector<CsPlan> plans;
};

// SGM_CS: bit 0 stride-1 convs, bit 1 strided down convs (fused with the residual branch), bit 2 transposed convs
static int cs_mask() {
  static int v = -1;
  if (v < 0) {
    v = 7;
    if (const char* env = getenv("SGM_CS")) v = atoi(env);
  }
  return v;
}

int cs_pack(const sgm_conv_desc* m, const sgm_conv_desc* second, TcConv* c) {
  c->cs_state = nullptr;
  if (c->flat0 || m->kernel != 3 || !tma_available()) return SGM_OK;   // 3-D k3 convs only
  const int bit = c->mode == MODE_S1 ? 1 : (c->mode == MODE_S2 ? 2 : 4);
  if (!(cs_mask() & bit)) return SGM_OK;
  // Which layers: every strided block (>= 16 input channels), stride-1 convs with >= 64 input channels, transposed
  // convs with >= 64.  Stride-1 layers with <= 32 channels and the transposed convs with <= 16 output channels stay on
  // the plane-sweep kernels.  (The strided 16-channel block, K = 27 blocks only, is bound by the strided TMA loads of
  // its parity slabs and by the epilogue: 0.65 -> 0.27 ms per 125 windows against the one-CTA-per-brick kernel.)
  if (c->mode == MODE_S1 && m->cin < 64) return SGM_OK;
  static const int s2_min = getenv("SGM_CS_S2MIN") ? atoi(getenv("SGM_CS_S2MIN")) : 16;
  if (c->mode == MODE_S2 && m->cin < s2_min) return SGM_OK;
  if (c->mode == MODE_T2 && (m->cin < 64 || c->ntot % 16 != 0)) return SGM_OK;
  auto* st = new CsState();
  CsPack& p = st->pack;
  p.nkc = c->cgin / 2;
  static const int n_pref = getenv("SGM_CS_N") ? atoi(getenv("SGM_CS_N")) : 128;  // measured: N = 128 beats 64 on every layer with >= 128 outputs
  if (c->mode == MODE_T2) {
    // parity classes folded into the MMA N dimension: 8 classes x 16 output channels, one tap block per input shift
    p.N = 16, p.NB = 128;
    for (int s0 = 0; s0 < 2; ++s0)
      for (int s1 = 0; s1 < 2; ++s1)
        for (int s2 = 0; s2 < 2; ++s2) p.taps.push_back({{s0, s1, s2}, 0, p.taps.empty() ? 1 : 0});
  } else {
    p.N = (c->ntot % n_pref == 0) ? n_pref : (c->ntot % 64 == 0 ? 64 : (c->ntot % 32 == 0 ? 32 : 16));
    p.N = std::min(p.N, c->ntot);
    p.NB = p.N;
    for (int k0 = 0; k0 < 3; ++k0)
      for (int k1 = 0; k1 < 3; ++k1)
        for (int k2 = 0; k2 < 3; ++k2) {
          CsTap t;
          const int kk[3] = {k0, k1, k2};
          int slab = 0;
          for (int a = 0; a < 3; ++a) {
            if (c->mode == MODE_S1) {
              t.shift[a] = kk[a] - 1;
            } else {  // stride 2: input 2o + k - 1 = 2 (o + e) + r: parity slab bit r, shift e
              const int r = kk[a] == 1 ? 0 : 1, e = kk[a] == 0 ? -1 : 0;
              t.shift[a] = e;
              slab |= r << (2 - a);
            }
          }
          t.slab = slab;
          t.first = p.taps.empty() ? 1 : 0;
          p.taps.push_back(t);
        }
  }
  p.ntap = (int)p.taps.size();
  p.ncoblk = c->ntot / p.N;
  const int nA = c->segA_cg * 8;
  const int ntaps = 27;
  auto wsrc = [&](int fco, int ci, int tap) -> float {
    const sgm_conv_desc* src = fco < nA ? m : second;
    const int co = fco < nA ? fco : fco - nA;
    if (!src || co >= src->cout || ci >= m->cin) return 0.f;
    if (src->kind == SGM_KIND_CONV_TRANSPOSE) return src->weight[((size_t)ci * src->cout + co) * ntaps + tap];
    return src->weight[((size_t)co * src->cin + ci) * ntaps + tap];
  };
  // [coblk][chunk][tap block][k half 2][NB][8] bf16
  std::vector<uint16_t> w((size_t)p.ncoblk * p.nkc * p.ntap * 2 * p.NB * 8, 0);
  for (int cb = 0; cb < p.ncoblk; ++cb)
    for (int kc = 0; kc < p.nkc; ++kc)
      for (int ti = 0; ti < p.ntap; ++ti) {
        const CsTap& t = p.taps[ti];
        for (int half = 0; half < 2; ++half)
          for (int row = 0; row < p.NB; ++row) {
            int fco, tap;
            if (c->mode == MODE_T2) {
              // class (p0, p1, p2) reads input shift s with tap: p = 0 -> (s = 0, k = 1); p = 1 -> (s = 0, k = 2), (s = 1, k = 0)
              const int cls = row / p.N;
              fco = cb * p.N + row % p.N;
              int kk[3];
              bool used = true;
              for (int a = 0; a < 3; ++a) {
                const int pbit = (cls >> (2 - a)) & 1;
                if (pbit == 0) kk[a] = 1, used = used && t.shift[a] == 0;
                else kk[a] = t.shift[a] == 0 ? 2 : 0;
              }
              if (!used) continue;
              tap = (kk[0] * 3 + kk[1]) * 3 + kk[2];
            } else {
              fco = cb * p.N + row;
              int kk[3];
              for (int a = 0; a < 3; ++a) {
                if (c->mode == MODE_S1) kk[a] = t.shift[a] + 1;
                else kk[a] = ((t.slab >> (2 - a)) & 1) == 0 ? 1 : (t.shift[a] == -1 ? 0 : 2);
              }
              tap = (kk[0] * 3 + kk[1]) * 3 + kk[2];
              if (m->kind == SGM_KIND_CONV_TRANSPOSE) tap = ntaps - 1 - tap;  // stride-1 transposed conv == flipped conv
            }
            for (int k8 = 0; k8 < 8; ++k8) {
              const int ci = (kc * 2 + half) * 8 + k8;
              w[(((((size_t)cb * p.nkc + kc) * p.ntap + ti) * 2 + half) * p.NB + row) * 8 + k8] = f2bf(wsrc(fco, ci, tap));
            }
          }
      }
  if (cudaMalloc(&c->cs_w, w.size() * 2) != cudaSuccess) {
    delete st;
    set_error("cs_pack: cudaMalloc failed: %s", cudaGetErrorString(cudaGetLastError()));
    return SGM_ERR_CUDA;
  }
  cudaMemcpy(c->cs_w, w.data(), w.size() * 2, cudaMemcpyHostToDevice);
  c->cs_state = st;
  return SGM_OK;
}

void cs_free(TcConv* c) {
  if (c->cs_w) cudaFree(c->cs_w);
  c->cs_w = nullptr;
  delete reinterpret_cast<CsState*>(c->cs_state);
  c->cs_state = nullptr;
}

bool cs_applicable(const TcConv& c, const TcIO& io) {
  if (!c.cs_state || io.out_kind != OUT_CG8) return false;
  if (io.cg0 % 2 != 0) return false;  // a 16-channel chunk never straddles the two concatenated inputs
  return true;
}

static int cs_plan(const TcConv& c, const TcIO& io, const CsPack& p, CsPlan& pl) {
  CsArgs& a = pl.args;
  memset(&a, 0, sizeof(a));
  a.mode = c.mode;
  a.cg0 = io.cg0, a.cg1 = io.cg1, a.nkc = p.nkc;
  for (int i = 0; i < 3; ++i) {
    a.id[i] = io.id[i], a.od[i] = io.od[i];
    a.rd[i] = c.mode == MODE_T2 ? io.id[i] : io.od[i];
  }
  a.N = p.N, a.NB = p.NB, a.ncls = c.mode == MODE_T2 ? 8 : 1, a.ntap = p.ntap, a.ncoblk = p.ncoblk;
  a.nwin = io.n;
  int addH[3];
  for (int i = 0; i < 3; ++i) {
    if (c.mode == MODE_S1) a.par[i] = 1, a.lo[i] = 1, addH[i] = 2;
    else if (c.mode == MODE_S2) a.par[i] = 2, a.lo[i] = 1, addH[i] = 1;
    else a.par[i] = 1, a.lo[i] = 0, addH[i] = 1;
  }
  a.nslab = c.mode == MODE_S2 ? 8 : 1;
  const int cols_tile = a.ncls * a.N;
  a.Gw = std::max(1, std::min(a.ntap, 18432 / (a.NB * 32)));
  a.ngw = ceil_div(a.ntap, a.Gw);
  a.w_stage_bytes = round_up(a.Gw * a.NB * 32, 128);
  a.wstages = 4;
  const int misc = (4 * kMaxStages + 4) * 8 + 16 + 128;
  const int nsm = sm_count();
  static const int cand[] = {1, 2, 3, 4, 6, 8, 12, 16, 24, 32, 48};
  static const int gcand[] = {1, 2, 3, 4, 6, 8};
  static const int force_g = getenv("SGM_CS_G") ? atoi(getenv("SGM_CS_G")) : 0;
  double best = 1e30;
  int bt[3] = {0, 0, 0}, bG = 0, bst = 0;
  for (int c0 : cand)
    for (int c1 : cand)
      for (int c2 : cand) {
        const int t[3] = {std::min(c0, a.rd[0]), std::min(c1, a.rd[1]), std::min(c2, a.rd[2])};
        const int H[3] = {t[0] + addH[0], t[1] + addH[1], t[2] + addH[2]};
        const int P = round_up(H[0] * H[1] * H[2], 8);
        if (c.mode == MODE_S2 ? (H[2] * 2 > 256 || H[1] * 2 > 256 || H[0] * 2 > 256) : (H[2] * 8 > 256 || H[1] > 256 || H[0] > 256))
          continue;  // TMA box limits
        const int rf = (a.lo[0] * H[1] + a.lo[1]) * H[2] + a.lo[2];
        const int rl = ((a.lo[0] + t[0] - 1) * H[1] + a.lo[1] + t[1] - 1) * H[2] + a.lo[2] + t[2] - 1;
        const int ntl = ceil_div(rl - rf + 1, 128);
        const long long nbricks = (long long)ceil_div(a.rd[0], t[0]) * ceil_div(a.rd[1], t[1]) * ceil_div(a.rd[2], t[2]);
        for (int G : gcand) {
          if (force_g && G != force_g) continue;
          // Several windows per unit (G > 1) are implemented (shared weight stages for 2-4 small windows) but no
          // launch of the benched or tested shapes ever selected them: opt-in through SGM_CS_G until they have run.
          if (!force_g && G > 1) break;
          if (G > io.n) break;
          if (nbricks > 1 && G > 1) break;  // several windows per unit only when a window is a single brick
          const int cols = G * ntl * cols_tile;
          if (cols > 512 || G * ntl > kIssuers * kMaxTilesPerIssuer) break;
          const long long a_stage = (long long)G * a.nslab * 2 * P * 16;
          const long long room = kSmemMax - 1024 - misc - (long long)a.wstages * a.w_stage_bytes;
          const int ast = (int)std::min<long long>(4, room / a_stage);
          if (ast < 2) break;
          const int nbuf = cols <= 256 ? 2 : 1;
          const long long units = (long long)ceil_div(io.n, G) * nbricks * p.ncoblk;
          const double rounds = (double)((units + nsm - 1) / nsm);
          // the tensor pipe is shared by the issuers: a unit's MMAs take at least their summed pipe time
          const double pipe = (double)G * ntl * p.nkc * p.ntap * mma_clk(a.NB);
          const double epi = (double)ceil_div(G * ntl, 2) * a.ncls * (a.N / 16) * 1300.0;
          const double load = ((double)a_stage * p.nkc + (double)p.nkc * p.ntap * a.NB * 32) / 24.0;
          const double unit = nbuf == 2 ? std::max(std::max(pipe, epi), load) + 1500.0 : std::max(pipe, load) + epi + 1500.0;
          const double cost = rounds * unit;
          if (cost < best) best = cost, bt[0] = t[0], bt[1] = t[1], bt[2] = t[2], bG = G, bst = ast;
        }
      }
  SGM_REQUIRE(bG > 0, SGM_ERR_UNSUPPORTED, "cs_plan: no brick shape fits (cgin=%d, N=%d)", c.cgin, a.N);
  a.G = bG, a.astages = bst;
  for (int i = 0; i < 3; ++i) {
    a.t[i] = bt[i];
    a.H[i] = bt[i] + addH[i];
    a.nt[i] = ceil_div(a.rd[i], bt[i]);
  }
  a.P = round_up(a.H[0] * a.H[1] * a.H[2], 8);
  a.box_bytes = a.H[0] * a.H[1] * a.H[2] * 16;
  a.row_first = (a.lo[0] * a.H[1] + a.lo[1]) * a.H[2] + a.lo[2];
  const int row_last = ((a.lo[0] + a.t[0] - 1) * a.H[1] + a.lo[1] + a.t[1] - 1) * a.H[2] + a.lo[2] + a.t[2] - 1;
  a.ntiles = ceil_div(row_last - a.row_first + 1, 128);
  a.nbuf = a.G * a.ntiles * cols_tile <= 256 ? 2 : 1;
  a.a_g_units = a.nslab * 2 * a.P;
  a.a_stage_units = a.G * a.a_g_units;
  a.nbricks = a.nt[0] * a.nt[1] * a.nt[2];
  a.nunits = ceil_div(io.n, a.G) * a.nbricks * p.ncoblk;
  SGM_REQUIRE(p.ntap <= kMaxTaps, SGM_ERR_UNSUPPORTED, "cs_plan: too many tap blocks");
  for (int ti = 0; ti < p.ntap; ++ti) {
    const CsTap& t = p.taps[ti];
    const int off = a.row_first + t.shift[0] * a.H[1] * a.H[2] + t.shift[1] * a.H[2] + t.shift[2] + t.slab * 2 * a.P;
    SGM_REQUIRE(off >= 0, SGM_ERR_INVALID, "cs_plan: negative tap offset");
    a.tap_a[ti] = (uint32_t)off;
    a.tap_c[ti] = (t.first ? 0x80000000u : 0u);
  }
  pl.smem_bytes = a.astages * a.a_stage_units * 16 + a.wstages * a.w_stage_bytes + misc;
  pl.grid = std::min(a.nunits, nsm);
  return SGM_OK;
}

int cs_launch(const TcConv& c, const TcIO& io, int* error_flag_dev, cudaStream_t st) {
  auto* state = reinterpret_cast<CsState*>(c.cs_state);
  const int key[8] = {io.id[0], io.id[1], io.id[2], io.od[0], io.od[1], io.od[2], io.n, io.cg0};
  const CsPlan* pe = nullptr;
  for (auto& e : state->plans)
    if (memcmp(e.key, key, sizeof(key)) == 0) pe = &e;
  if (!pe) {
    CsPlan e;
    memcpy(e.key, key, sizeof(key));
    int rc = cs_plan(c, io, state->pack, e);
    if (rc) return rc;
    state->plans.push_back(e);
    pe = &state->plans.back();
  }
  CsArgs a = pe->args;
  a.w = c.cs_w, a.bias = c.bias;
  a.outA = (__nv_bfloat16*)io.outA, a.cgA = io.cgA, a.outB = (__nv_bfloat16*)io.outB, a.cgB = io.cgB;
  a.segA_cg = c.segA_cg, a.actA = c.actA, a.alphaA = c.alphaA;
  a.res = (const __nv_bfloat16*)io.res;
  a.error_flag = error_flag_dev;
  static const bool dbg = getenv("SGM_DEBUG") != nullptr;
  if (dbg)
    fprintf(stderr,
            "[cs_launch] mode=%d N=%d NB=%d ntap=%d nkc=%d ncoblk=%d rd=(%d,%d,%d) t=(%d,%d,%d) H=(%d,%d,%d) P=%d nslab=%d "
            "row_first=%d ntiles=%d G=%d nbuf=%d astages=%d a_stage=%d B wstages=%d Gw=%d w_stage=%d B smem=%d units=%d "
            "grid=%d\n",
            a.mode, a.N, a.NB, a.ntap, a.nkc, a.ncoblk, a.rd[0], a.rd[1], a.rd[2], a.t[0], a.t[1], a.t[2], a.H[0], a.H[1],
            a.H[2], a.P, a.nslab, a.row_first, a.ntiles, a.G, a.nbuf, a.astages, a.a_stage_units * 16, a.wstages, a.Gw,
            a.w_stage_bytes, pe->smem_bytes, a.nunits, pe->grid);
  static DeviceOnce attr_once;
  if (attr_once.first()) {
    SGM_CUDA_CHECK(cudaFuncSetAttribute(cs_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax));
  }
  static const bool trace_on = getenv("SGM_TRACE") != nullptr;
  static long long* trace_dev = nullptr;
  if (trace_on) {
    if (!trace_dev) cudaMalloc(&trace_dev, 32 * sizeof(long long));
    cudaMemsetAsync(trace_dev, 0, 32 * sizeof(long long), st);
    a.trace = trace_dev;
  }
  CUtensorMap tm0, tm1;
  memset(&tm0, 0, sizeof(tm0));
  memset(&tm1, 0, sizeof(tm1));
  int rc = make_brick_map(&tm0, io.in0, io.n * io.cg0, io.id, a.H, c.mode == MODE_S2 ? a.par : nullptr);
  if (!rc && io.in1) rc = make_brick_map(&tm1, io.in1, io.n * io.cg1, io.id, a.H);
  if (rc) return rc;
  cs_conv_kernel<<<pe->grid, kThreads, pe->smem_bytes, st>>>(a, tm0, tm1);
  SGM_CUDA_CHECK(cudaGetLastError());
  if (trace_on) {
    long long t[32];
    cudaStreamSynchronize(st);
    cudaMemcpy(t, trace_dev, sizeof(t), cudaMemcpyDeviceToHost);
    auto d = [&](int i) { return t[i] ? (double)(t[i] - t[0]) : -1.0; };
    fprintf(stderr,
            "[cs trace] mode=%d N=%d ntap=%d nkc=%d units=%d grid=%d G=%d ntiles=%d nbuf=%d | issued %.0f %.0f %.0f %.0f | "
            "accdone %.0f %.0f %.0f %.0f | epidone %.0f %.0f %.0f %.0f | end %.0f cycles\n",
            a.mode, a.N, a.ntap, a.nkc, a.nunits, pe->grid, a.G, a.ntiles, a.nbuf, d(1), d(2), d(3), d(4), d(8), d(9), d(10),
            d(11), d(14), d(15), d(16), d(17), d(20));
  }
  return SGM_OK;
}

}  // namespace tc
}  // namespace sgm
