// "N-fold" tcgen05 implicit-GEMM convolution: stride-1 3x3x3 convs with few output channels (the
// residual units at the two finest resolutions and the head: 70 % of the UNet's FLOPs have Cout <= 32).
//
// Why.  An SS-mode tcgen05.mma (M=128, K=16) reads its 4 KB A tile from shared memory in ~32 clk, while
// the tensor pipe needs only N/2 clk: with N = Cout = 16 the plain one-MMA-per-tap formulation
// (conv_tc.cu) is bound by shared-memory operand bandwidth at 25 % of the tensor peak.  Here the three
// d2 taps are folded into N: B = [W(k0,k1,0) | W(k0,k1,1) | W(k0,k1,2)]  (N' = 3*CS columns), so one A
// tile feeds three taps -- 9 MMAs per K block instead of 27, one third of the A traffic:
//     P_k2[p] = sum_{k0,k1,ci} W[k0,k1,k2][co][ci] * X[ci][p + (k0-1)*H1*H2 + (k1-1)*H2]       (tensor core)
//     out[q]  = P_0[q-1] + P_1[q] + P_2[q+1]                                                  (epilogue)
// with p, q padded-linear positions of the shared-memory halo brick, as in conv_tc.cu.
//
// The d2 shift of the epilogue is a lane shift in TMEM (lane = GEMM row = brick position).  Inside a warp it
// is two shuffles per channel; a warp may only read its own 32-lane TMEM quarter, so the first / last lane of
// each warp gets its outer neighbour through a 2*CS-float shared-memory exchange between the four warps of
// an epilogue group (one named barrier per tile).  Consecutive tiles overlap by two rows, so rows 0 and 127
// of a tile only serve as neighbours: one tile = 126 outputs, 98 % of the MMA rows are useful.
// (A first version made every 8-row core matrix overlap the previous one by two rows with SBO = 96 B, which
// needs no exchange at all -- and measured 96 clk per MMA instead of 32: the operand fetch wants core
// matrices on 128-byte lines.  profiles/r01_nf_notes.md.)
//
// Persistent, one CTA per SM (320 threads): warps 0-7 epilogue (two groups of four alternate tiles),
// warp 8 = TMA producer (double-buffered halo bricks, weights resident in shared memory for the whole
// kernel), warp 9 = one elected lane issues the MMAs into a ring of 4-8 TMEM accumulator slots.
// Identity residuals (up-path units, head) are read from the brick centre in shared memory.
// All mbarrier waits are bounded (error flag + trap) so a protocol bug cannot hang the GPU.
#include "conv_tc.cuh"
#include "tc_ptx.cuh"

#include <algorithm>
#include <stdlib.h>
#include <string.h>

namespace sgm {
namespace tc {

namespace {
using namespace tcptx;

constexpr int kNfThreads = 320;
constexpr int kNfSmemMax = 227 * 1024;
constexpr int kTileOut = 126;  // outputs per M=128 tile (rows 1..126; consecutive tiles overlap by two rows)
constexpr int kSlack = 136;    // rows the last tile may read past the last brick position

struct NfArgs {
  int D[3];                 // tensor extents (input == output)
  int t[3], H[3], nt[3];
  int H12, P, Ppad;         // brick positions per channel group, padded stride (16-byte units)
  uint32_t m12, m2;         // ceil(2^32 / H12), ceil(2^32 / H2): exact division of positions < 2^16
  int row_first, ntiles;
  int nbricks, bricks_per_win;
  int nslot_log2, slot_stride;
  int cgA, c_real, act;
  float alpha;
  int res_mode;             // 0 none, 1 global CG8 tensor, 2 identity (centre of the brick)
  int out_kind;             // OUT_CG8 | OUT_PLANAR | OUT_BLEND
  int pl_weighted;
  int ad0, ad1, ad2, wo[3]; // OUT_BLEND: accumulator extent and window origin
  const __nv_bfloat16* w;
  const float* bias;
  __nv_bfloat16* out;
  const __nv_bfloat16* res;
  float* pl_out;
  long long pl_cstride, pl_nstride;
  const float* imap0;
  const float* imap1;
  const float* imap2;
  float imap_floor;
  int* error_flag;
  long long* trace;
};

__device__ __forceinline__ void ld_cols(uint32_t taddr, uint32_t* v, int nreal) {
  if (nreal >= 5) tc_ld_x8(taddr, v);
  else if (nreal >= 3) tc_ld_x4(taddr, v);
  else if (nreal == 2) tc_ld_x2(taddr, v);
  else tc_ld_x1(taddr, v);
}

template <int CS, int NCGP>
__global__ void __launch_bounds__(kNfThreads, 1)
nf_conv_kernel(const NfArgs a, const __grid_constant__ CUtensorMap tmap) {
  constexpr int NP = (3 * CS + 15) / 16 * 16;  // MMA N
  constexpr int NKB = 9 * NCGP;                // K blocks (16 input channels x one (k0,k1) tap pair)
  constexpr int CG = 2 * NCGP;                 // input channel groups
  constexpr uint32_t W_BYTES = NKB * NP * 32;
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  uint8_t* w_smem = smem;
  uint8_t* a_smem = smem + W_BYTES;
  const uint32_t brick_bytes = (uint32_t)CG * a.Ppad * 16u;
  float* bias_s = reinterpret_cast<float*>(a_smem + 2 * brick_bytes);
  float* xchg = bias_s + 32;  // [group 2][parity 2][warp 4][left/right 2][CS]
  uint64_t* bars = reinterpret_cast<uint64_t*>(xchg + 32 * CS);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 21);
  const uint32_t bar0 = smem_u32(bars);
  const uint32_t WBAR = bar0;
  auto AFULL = [&](int b) { return bar0 + 8u * (1 + b); };
  auto AEMPTY = [&](int b) { return bar0 + 8u * (3 + b); };
  auto TFULL = [&](int s) { return bar0 + 8u * (5 + s); };
  auto TEMPTY = [&](int s) { return bar0 + 8u * (13 + s); };
  const int nslot = 1 << a.nslot_log2;
  const bool tr = a.trace != nullptr && blockIdx.x == 0;

  if (tid == 0) {
    mbar_init(WBAR, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(AFULL(b), 1);
      mbar_init(AEMPTY(b), 1 + 8);  // MMA commit + one arrival per epilogue warp
    }
    for (int s = 0; s < 8; ++s) {
      mbar_init(TFULL(s), 1);
      mbar_init(TEMPTY(s), 4);      // the four warps of the epilogue group that owns the slot
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (tr) a.trace[0] = clock64();
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // zero both brick buffers once: TMA later rewrites positions [0, P) of every channel group, the padding
  // rows (read by the last tiles, results discarded) stay finite
  for (uint32_t i = tid; i < 2 * brick_bytes / 16; i += kNfThreads)
    reinterpret_cast<uint4*>(a_smem)[i] = make_uint4(0, 0, 0, 0);
  if (tid < 32) bias_s[tid] = tid < CS ? __ldg(a.bias + tid) : 0.f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 8) {
    // ============================ producer: weights once, then the halo bricks ============================
    if (lane == 0) {
      mbar_expect_tx(WBAR, W_BYTES);
      bulk_g2s(smem_u32(w_smem), a.w, W_BYTES, WBAR);
      const uint32_t a_base = smem_u32(a_smem);
      int it = 0;
      for (int brick = blockIdx.x; brick < a.nbricks; brick += gridDim.x, ++it) {
        const int buf = it & 1;
        if (it >= 2) mbar_wait_or_trap(AEMPTY(buf), (uint32_t)((it >> 1) - 1) & 1u, a.error_flag, 11);
        const int n = brick / a.bricks_per_win;
        int r = brick - n * a.bricks_per_win;
        const int b2 = r % a.nt[2];
        r /= a.nt[2];
        const int b1 = r % a.nt[1], b0 = r / a.nt[1];
        mbar_expect_tx(AFULL(buf), (uint32_t)(CG * a.P * 16));
        // one 4-D box {H2*8 elements, H1, H0, 1 group}; coordinates outside the window are zero-filled
        // by the hardware == the conv's zero padding (MONAI convolves every window in isolation)
#pragma unroll
        for (int cg = 0; cg < CG; ++cg)
          tma_load_4d(a_base + buf * brick_bytes + (uint32_t)(cg * a.Ppad) * 16u, &tmap, (b2 * a.t[2] - 1) * 8,
                      b1 * a.t[1] - 1, b0 * a.t[0] - 1, n * CG + cg, AFULL(buf));
      }
    }
  } else if (warp == 9) {
    // ============================ MMA issuer: one elected lane ============================
    if (elect_one()) {
      // instruction descriptor: D=f32, A=B=bf16, both K-major, N>>3 at [17,23), M>>4 at [24,29)
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NP >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t a_base16 = smem_u32(a_smem) >> 4;
      const uint32_t w_base16 = smem_u32(w_smem) >> 4;
      const uint32_t a_hi = 8u | (1u << 14);   // SBO = 128 B: 128 consecutive brick positions
      const uint32_t b_hi = 8u | (1u << 14);   // SBO = 128 B
      const uint32_t a_lbo = ((uint32_t)a.Ppad & 0x3FFFu) << 16;  // next 8 input channels: next channel group
      const uint32_t b_lbo = ((uint32_t)NP & 0x3FFFu) << 16;      // next 8 input channels of the filter block
      mbar_wait_or_trap(WBAR, 0u, a.error_flag, 12);
      int tile_ctr = 0, it = 0;
      for (int brick = blockIdx.x; brick < a.nbricks; brick += gridDim.x, ++it) {
        const int buf = it & 1;
        mbar_wait_or_trap(AFULL(buf), (uint32_t)(it >> 1) & 1u, a.error_flag, 13);
        tc_fence_after();
        if (tr && it < 4) a.trace[1 + it] = clock64();
        const uint32_t abuf16 = a_base16 + (uint32_t)buf * (brick_bytes >> 4);
        for (int t = 0; t < a.ntiles; ++t, ++tile_ctr) {
          const int slot = tile_ctr & (nslot - 1);
          const int use = tile_ctr >> a.nslot_log2;
          if (use > 0) {
            mbar_wait_or_trap(TEMPTY(slot), (uint32_t)(use - 1) & 1u, a.error_flag, 14);
            tc_fence_after();
          }
          const uint32_t dcol = tmem_base + (uint32_t)(slot * a.slot_stride);
          const uint32_t base = abuf16 + (uint32_t)(a.row_first - 1 + kTileOut * t);
#pragma unroll
          for (int kb = 0; kb < NKB; ++kb) {
            const int k0 = kb / (3 * NCGP), k1 = (kb / NCGP) % 3, cp = kb % NCGP;
            const uint32_t a_lo = ((base + (uint32_t)(cp * 2 * a.Ppad + (k0 - 1) * a.H12 + (k1 - 1) * a.H[2])) & 0x3FFFu) | a_lbo;
            const uint32_t b_lo = ((w_base16 + (uint32_t)(kb * NP * 2)) & 0x3FFFu) | b_lbo;
            tc_mma(dcol, ((uint64_t)a_hi << 32) | a_lo, ((uint64_t)b_hi << 32) | b_lo, idesc, kb > 0 ? 1u : 0u);
          }
          tc_commit(TFULL(slot));
        }
        tc_commit(AEMPTY(buf));  // every MMA that reads this brick has completed when this arrives
        if (tr && it < 4) a.trace[5 + it] = clock64();
      }
    }
    __syncwarp();
  } else {
    // ============================ epilogue warps 0..7 ============================
    const int egroup = warp >> 2, quarter = warp & 3;
    const int row = quarter * 32 + lane;  // GEMM row of this lane == TMEM lane
    const bool rvalid = row >= 1 && row <= kTileOut;
    const long long vox = (long long)a.D[0] * a.D[1] * a.D[2];
    const uint32_t tlane = tmem_base + ((uint32_t)(quarter * 32) << 16);
    int tile_ctr = 0, it = 0, gtile = 0;
    for (int brick = blockIdx.x; brick < a.nbricks; brick += gridDim.x, ++it) {
      const int buf = it & 1;
      const int n = brick / a.bricks_per_win;
      int r = brick - n * a.bricks_per_win;
      const int b2 = r % a.nt[2];
      r /= a.nt[2];
      const int b1 = r % a.nt[1], b0 = r / a.nt[1];
      const int org0 = b0 * a.t[0] - 1, org1 = b1 * a.t[1] - 1, org2 = b2 * a.t[2] - 1;  // brick origin
      const uint8_t* brick_s = a_smem + (size_t)buf * brick_bytes;
      // Observe the TMA barrier of this brick: (1) identity residuals are read from the brick with generic
      // loads, (2) it keeps a warp without tiles in a brick from running ahead and arriving twice on AEMPTY.
      if (!mbar_wait(AFULL(buf), (uint32_t)(it >> 1) & 1u, a.error_flag, 15)) break;
      bool ok = true;
      for (int t = 0; t < a.ntiles; ++t, ++tile_ctr) {
        if ((tile_ctr & 1) != egroup) continue;
        const int slot = tile_ctr & (nslot - 1);
        const int use = tile_ctr >> a.nslot_log2;
        // ---- geometry of this lane's row (before waiting for the tensor core)
        const int q = a.row_first - 1 + kTileOut * t + row;
        const int h0 = (int)__umulhi((uint32_t)q, a.m12);
        const int q12 = q - h0 * a.H12;
        const int h1 = (int)__umulhi((uint32_t)q12, a.m2);
        const int h2 = q12 - h1 * a.H[2];
        const int r0 = org0 + h0, r1 = org1 + h1, r2 = org2 + h2;
        const bool valid = rvalid && h0 >= 1 && h0 <= a.t[0] && h1 >= 1 && h1 <= a.t[1] && h2 >= 1 && h2 <= a.t[2] &&
                           r0 < a.D[0] && r1 < a.D[1] && r2 < a.D[2];
        const int opos = (r0 * a.D[1] + r1) * a.D[2] + r2;
        float imw = 1.f;
        if ((a.pl_weighted || a.out_kind == OUT_BLEND) && valid)
          imw = fmaxf(__fmul_rn(__fmul_rn(__ldg(a.imap0 + r0), __ldg(a.imap1 + r1)), __ldg(a.imap2 + r2)), a.imap_floor);
        uint4 gres[CS / 8];
        if (a.res_mode == 1) {
#pragma unroll
          for (int pc = 0; pc < CS / 8; ++pc) {
            gres[pc] = make_uint4(0, 0, 0, 0);
            if (valid && pc < a.cgA)
              gres[pc] = __ldg(reinterpret_cast<const uint4*>(a.res + (((long long)n * a.cgA + pc) * vox + opos) * 8));
          }
        }
        ok = mbar_wait(TFULL(slot), (uint32_t)use & 1u, a.error_flag, 16);
        if (!ok) break;
        tc_fence_after();
        // ---- the whole accumulator row of this lane -> registers, then the TMEM slot goes back to the MMA warp
        const uint32_t tcol = tlane + (uint32_t)(slot * a.slot_stride);
        uint32_t p0[CS], p1[CS], p2[CS];
#pragma unroll
        for (int c = 0; c < CS; ++c) p0[c] = 0u, p1[c] = 0u, p2[c] = 0u;
#pragma unroll
        for (int pc = 0; pc < CS / 8; ++pc) {
          const int nreal = a.c_real - 8 * pc;  // uniform
          if (nreal > 0) {
            ld_cols(tcol + 8 * pc, p0 + 8 * pc, nreal);
            ld_cols(tcol + CS + 8 * pc, p1 + 8 * pc, nreal);
            ld_cols(tcol + 2 * CS + 8 * pc, p2 + 8 * pc, nreal);
          }
        }
        tc_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(TEMPTY(slot));
        // ---- outer neighbours of the warp's first / last lane come from the adjacent warps of the group
        float* xg = xchg + ((egroup * 2 + (gtile & 1)) * 4) * 2 * CS;
        float* xw = xg + quarter * 2 * CS;
        if (lane == 31) {
#pragma unroll
          for (int c4 = 0; c4 < CS / 4; ++c4)
            reinterpret_cast<uint4*>(xw)[c4] = make_uint4(p0[4 * c4], p0[4 * c4 + 1], p0[4 * c4 + 2], p0[4 * c4 + 3]);
        }
        if (lane == 0) {
#pragma unroll
          for (int c4 = 0; c4 < CS / 4; ++c4)
            reinterpret_cast<uint4*>(xw + CS)[c4] = make_uint4(p2[4 * c4], p2[4 * c4 + 1], p2[4 * c4 + 2], p2[4 * c4 + 3]);
        }
        asm volatile("bar.sync %0, 128;" ::"r"(1 + egroup) : "memory");
        ++gtile;
        const float* xl = xg + ((quarter + 3) & 3) * 2 * CS;       // previous warp: P0 of its lane 31
        const float* xr = xg + ((quarter + 1) & 3) * 2 * CS + CS;  // next warp: P2 of its lane 0
#pragma unroll
        for (int pc = 0; pc < CS / 8; ++pc) {
          const int nreal = a.c_real - 8 * pc;  // uniform
          if (nreal <= 0) break;
          float v[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            float left = __shfl_up_sync(0xffffffffu, __uint_as_float(p0[8 * pc + c]), 1);
            float right = __shfl_down_sync(0xffffffffu, __uint_as_float(p2[8 * pc + c]), 1);
            if (lane == 0) left = xl[8 * pc + c];
            if (lane == 31) right = xr[8 * pc + c];
            float x = (left + __uint_as_float(p1[8 * pc + c])) + right + bias_s[8 * pc + c];
            if (a.act) x = prelu(x, a.alpha);
            v[c] = x;
          }
          if (valid) {
            if (a.res_mode == 1) {
              float rr[8];
              unpack8(gres[pc], rr);
#pragma unroll
              for (int c = 0; c < 8; ++c) v[c] += rr[c];
            } else if (a.res_mode == 2) {
              float rr[8];
              unpack8(*reinterpret_cast<const uint4*>(brick_s + ((size_t)pc * a.Ppad + q) * 16), rr);
#pragma unroll
              for (int c = 0; c < 8; ++c) v[c] += rr[c];
            }
            if (a.out_kind == OUT_CG8) {
              if (pc < a.cgA)
                *reinterpret_cast<uint4*>(a.out + (((long long)n * a.cgA + pc) * vox + opos) * 8) = pack8(v);
            } else if (a.out_kind == OUT_PLANAR) {
              float* dst = a.pl_out + (long long)n * a.pl_nstride + (long long)(8 * pc) * a.pl_cstride + opos;
#pragma unroll
              for (int c = 0; c < 8; ++c)
                if (c < nreal) __stcs(dst + c * a.pl_cstride, a.pl_weighted ? __fmul_rn(v[c], imw) : v[c]);
            } else {  // OUT_BLEND: acc += seg * w for this window (read-modify-write form; one window per launch)
              const int g0 = a.wo[0] + r0;
              if (g0 >= 0 && g0 < a.ad0) {
                float* dst = a.pl_out + (long long)(8 * pc) * a.pl_cstride +
                             ((long long)g0 * a.ad1 + (a.wo[1] + r1)) * a.ad2 + (a.wo[2] + r2);
                float oldv[8];
#pragma unroll
                for (int c = 0; c < 8; ++c)
                  if (c < nreal) oldv[c] = __ldcg(dst + c * a.pl_cstride);
#pragma unroll
                for (int c = 0; c < 8; ++c)  // seg *= w; out += seg (two roundings, as MONAI)
                  if (c < nreal) __stcg(dst + c * a.pl_cstride, __fadd_rn(oldv[c], __fmul_rn(v[c], imw)));
              }
            }
          }
        }
      }
      if (!ok) break;
      __syncwarp();
      if (lane == 0) mbar_arrive(AEMPTY(buf));  // this warp no longer reads the brick
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
  if (tr && tid == 0) a.trace[9] = clock64();
}

inline uint16_t f2bf(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7f800000u) != 0x7f800000u) u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}

inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

struct NfPlan {
  int key[5];
  NfArgs args;
  int smem_bytes, grid;
};

int nf_np(int cs) { return round_up(3 * cs, 16); }

int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

int nf_plan(const TcConv& c, const TcIO& io, NfPlan& pl) {
  NfArgs& a = pl.args;
  memset(&a, 0, sizeof(a));
  const int CG = 2 * c.nf_ncgp, NP = nf_np(c.nf_cs), NKB = 9 * c.nf_ncgp;
  const int w_bytes = NKB * NP * 32;
  const int fixed = w_bytes + 128 + 32 * c.nf_cs * 4 + 21 * 8 + 16 + 128;
  for (int i = 0; i < 3; ++i) a.D[i] = io.od[i];
  static const int cand[] = {1, 2, 3, 4, 6, 8, 12, 16, 24, 30, 32, 48, 64, 96};
  const int nsm = sm_count();
  const double t_mma = NKB * std::max(34.0, NP / 2.0);
  const double t_epi = 3.0 * round_up(c.c_real, 8) * 8.0 + 60.0;  // TMEM read at ~64 B/clk + fixed cost
  const double t_tile = std::max(t_mma, t_epi);
  double best = 1e30;
  int bt[3] = {0, 0, 0};
  for (int c0 : cand)
    for (int c1 : cand)
      for (int c2 : cand) {
        const int t[3] = {std::min(c0, a.D[0]), std::min(c1, a.D[1]), std::min(c2, a.D[2])};
        const int H[3] = {t[0] + 2, t[1] + 2, t[2] + 2};
        if (H[2] * 8 > 256 || H[1] > 256 || H[0] > 256) continue;  // TMA box limits
        const int P = H[0] * H[1] * H[2];
        const int Ppad = round_up(P + kSlack, 8);
        if (Ppad > 16383 || P >= 65536) continue;
        const long long smem = (long long)fixed + 2LL * CG * Ppad * 16;
        if (smem > kNfSmemMax) continue;
        const int span = (t[0] - 1) * H[1] * H[2] + (t[1] - 1) * H[2] + t[2];
        const int ntl = ceil_div(span, kTileOut);
        const long long nb = (long long)ceil_div(a.D[0], t[0]) * ceil_div(a.D[1], t[1]) * ceil_div(a.D[2], t[2]) * io.n;
        const double load = (double)CG * P * 16 / 24.0;
        const double cta = std::max(ntl * t_tile, load) + 500.0;
        const double waves = (double)((nb + nsm - 1) / nsm);
        const double cost = waves * cta + load;  // the first brick of a CTA is not overlapped
        if (cost < best) best = cost, bt[0] = t[0], bt[1] = t[1], bt[2] = t[2];
      }
  SGM_REQUIRE(bt[0] > 0, SGM_ERR_UNSUPPORTED, "nf_launch: no brick shape fits shared memory");
  for (int i = 0; i < 3; ++i) a.t[i] = bt[i], a.H[i] = bt[i] + 2, a.nt[i] = ceil_div(a.D[i], bt[i]);
  a.H12 = a.H[1] * a.H[2];
  a.P = a.H[0] * a.H12;
  a.Ppad = round_up(a.P + kSlack, 8);
  a.m12 = (uint32_t)((0x100000000ULL + a.H12 - 1) / a.H12);
  a.m2 = (uint32_t)((0x100000000ULL + a.H[2] - 1) / a.H[2]);
  a.row_first = a.H12 + a.H[2] + 1;
  const int span = (a.t[0] - 1) * a.H12 + (a.t[1] - 1) * a.H[2] + a.t[2];
  a.ntiles = ceil_div(span, kTileOut);
  a.bricks_per_win = a.nt[0] * a.nt[1] * a.nt[2];
  a.nbricks = a.bricks_per_win * io.n;
  int stride = 32;
  while (stride < NP) stride <<= 1;
  a.slot_stride = stride;
  a.nslot_log2 = stride <= 64 ? 3 : 2;  // 8 x 64 or 4 x 128 columns
  pl.smem_bytes = fixed + 2 * CG * a.Ppad * 16;
  pl.grid = std::min(nsm, a.nbricks);
  return SGM_OK;
}

template <int CS, int NCGP>
int launch_t(const NfArgs& a, const CUtensorMap& tm, int grid, int smem, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    SGM_CUDA_CHECK(cudaFuncSetAttribute(nf_conv_kernel<CS, NCGP>, cudaFuncAttributeMaxDynamicSharedMemorySize, kNfSmemMax));
    attr_set = true;
  }
  nf_conv_kernel<CS, NCGP><<<grid, kNfThreads, smem, st>>>(a, tm);
  SGM_CUDA_CHECK(cudaGetLastError());
  return SGM_OK;
}

}  // namespace

int nf_pack(const sgm_conv_desc& d, TcConv* c) {
  c->nf_cs = 0;
  if (getenv("SGM_NO_NF")) return SGM_OK;
  if (d.kind != SGM_KIND_CONV || d.kernel != 3 || d.stride != 1 || c->flat0 || c->mode != MODE_S1) return SGM_OK;
  if (d.cout > 32 || d.cin > 32) return SGM_OK;
  const int cs = round_up(d.cout, 8);
  const int ncgp = cs <= 16 ? 1 : 2;
  if (c->cgin != 2 * ncgp) return SGM_OK;  // one instantiation per CS: the input must span exactly NCGP pairs
  const int NP = nf_np(cs), NKB = 9 * ncgp;
  std::vector<uint16_t> w((size_t)NKB * 2 * NP * 8, 0);
  for (int k0 = 0; k0 < 3; ++k0)
    for (int k1 = 0; k1 < 3; ++k1)
      for (int cp = 0; cp < ncgp; ++cp) {
        const int kb = (k0 * 3 + k1) * ncgp + cp;
        for (int kc = 0; kc < 2; ++kc)
          for (int k2 = 0; k2 < 3; ++k2)
            for (int co = 0; co < d.cout; ++co)
              for (int k8 = 0; k8 < 8; ++k8) {
                const int ci = (cp * 2 + kc) * 8 + k8;
                if (ci >= d.cin) continue;
                const int tap = (k0 * 3 + k1) * 3 + k2;
                w[(((size_t)kb * 2 + kc) * NP + (k2 * cs + co)) * 8 + k8] =
                    f2bf(d.weight[((size_t)co * d.cin + ci) * 27 + tap]);
              }
      }
  if (cudaMalloc(&c->nf_w, w.size() * 2) != cudaSuccess) {
    set_error("nf_pack: cudaMalloc failed: %s", cudaGetErrorString(cudaGetLastError()));
    return SGM_ERR_CUDA;
  }
  SGM_CUDA_CHECK(cudaMemcpy(c->nf_w, w.data(), w.size() * 2, cudaMemcpyHostToDevice));
  c->nf_cs = cs, c->nf_ncgp = ncgp;
  c->nf_plan_cache = new std::vector<NfPlan>();
  return SGM_OK;
}

void nf_free(TcConv* c) {
  if (c->nf_w) cudaFree(c->nf_w);
  c->nf_w = nullptr;
  delete reinterpret_cast<std::vector<NfPlan>*>(c->nf_plan_cache);
  c->nf_plan_cache = nullptr;
}

bool nf_applicable(const TcConv& c, const TcIO& io) {
  if (!c.nf_cs || !c.nf_w || !tma_available()) return false;
  if (io.in1 || io.cg1 || io.outB) return false;
  if (io.out_kind == OUT_BLEND && io.n != 1) return false;
  if (io.cg0 != 2 * c.nf_ncgp) return false;
  for (int i = 0; i < 3; ++i)
    if (io.id[i] != io.od[i]) return false;
  if (io.out_kind == OUT_CG8 && io.cgA * 8 < c.nf_cs) return false;
  if ((long long)io.od[0] * io.od[1] * io.od[2] >= (1LL << 31)) return false;
  return true;
}

int nf_launch(const TcConv& c, const TcIO& io, int* error_flag_dev, cudaStream_t st) {
  auto* plans = reinterpret_cast<std::vector<NfPlan>*>(c.nf_plan_cache);
  const int key[5] = {io.od[0], io.od[1], io.od[2], io.n, 0};
  const NfPlan* pe = nullptr;
  for (auto& e : *plans)
    if (memcmp(e.key, key, sizeof(key)) == 0) pe = &e;
  if (!pe) {
    NfPlan e;
    memcpy(e.key, key, sizeof(key));
    int rc = nf_plan(c, io, e);
    if (rc) return rc;
    plans->push_back(e);
    pe = &plans->back();
  }
  NfArgs a = pe->args;
  a.cgA = io.cgA, a.c_real = c.c_real, a.act = c.actA, a.alpha = c.alphaA;
  a.w = c.nf_w, a.bias = c.bias;
  a.out = (__nv_bfloat16*)io.outA, a.res = (const __nv_bfloat16*)io.res;
  a.res_mode = 0;
  if (io.res) a.res_mode = (io.res == io.in0 && io.cgA == 2 * c.nf_ncgp) ? 2 : 1;
  a.out_kind = io.out_kind, a.pl_weighted = io.pl_weighted;
  a.pl_out = io.pl_out, a.pl_cstride = io.pl_cstride, a.pl_nstride = io.pl_nstride;
  a.imap0 = io.imap[0], a.imap1 = io.imap[1], a.imap2 = io.imap[2], a.imap_floor = io.imap_floor;
  a.ad0 = io.ad0, a.ad1 = io.ad1, a.ad2 = io.ad2;
  for (int i = 0; i < 3; ++i) a.wo[i] = io.wo[i];
  a.error_flag = error_flag_dev;
  static const bool dbg = getenv("SGM_DEBUG") != nullptr;
  if (dbg)
    fprintf(stderr,
            "[nf_launch] CS=%d NCGP=%d D=(%d,%d,%d) n=%d t=(%d,%d,%d) P=%d Ppad=%d ntiles=%d nbricks=%d grid=%d smem=%d "
            "slots=%d x %d res=%d out=%d\n",
            c.nf_cs, c.nf_ncgp, a.D[0], a.D[1], a.D[2], io.n, a.t[0], a.t[1], a.t[2], a.P, a.Ppad, a.ntiles, a.nbricks,
            pe->grid, pe->smem_bytes, 1 << a.nslot_log2, a.slot_stride, a.res_mode, a.out_kind);
  static const bool trace_on = getenv("SGM_TRACE") != nullptr;
  static long long* trace_dev = nullptr;
  if (trace_on) {
    if (!trace_dev) cudaMalloc(&trace_dev, 16 * sizeof(long long));
    cudaMemsetAsync(trace_dev, 0, 16 * sizeof(long long), st);
    a.trace = trace_dev;
  }
  CUtensorMap tm;
  int rc = make_brick_map(&tm, io.in0, io.n * io.cg0, io.id, a.H);
  if (rc) return rc;
  if (c.nf_cs == 8) rc = launch_t<8, 1>(a, tm, pe->grid, pe->smem_bytes, st);
  else if (c.nf_cs == 16) rc = launch_t<16, 1>(a, tm, pe->grid, pe->smem_bytes, st);
  else if (c.nf_cs == 24) rc = launch_t<24, 2>(a, tm, pe->grid, pe->smem_bytes, st);
  else rc = launch_t<32, 2>(a, tm, pe->grid, pe->smem_bytes, st);
  if (rc) return rc;
  if (trace_on) {
    long long t[16];
    cudaStreamSynchronize(st);
    cudaMemcpy(t, trace_dev, sizeof(t), cudaMemcpyDeviceToHost);
    auto d = [&](int i) { return t[i] ? (double)(t[i] - t[0]) : -1.0; };
    fprintf(stderr,
            "[nf trace] CS=%d D=(%d,%d,%d) n=%d t=(%d,%d,%d) ntiles=%d bricks/cta=%.1f | brick landed %.0f %.0f %.0f %.0f | "
            "issued %.0f %.0f %.0f %.0f | end %.0f cycles\n",
            c.nf_cs, a.D[0], a.D[1], a.D[2], io.n, a.t[0], a.t[1], a.t[2], a.ntiles, (double)a.nbricks / pe->grid, d(1), d(2),
            d(3), d(4), d(5), d(6), d(7), d(8), d(9));
  }
  return SGM_OK;
}

}  // namespace tc
}  // namespace sgm
