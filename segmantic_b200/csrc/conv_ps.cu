// "Plane-sweep" tcgen05 implicit-GEMM convolution: stride-1 3x3x3 convs with few channels (the residual
// units at the two finest resolutions and the head -- 70 % of the UNet's FLOPs have Cout <= 32).
//
// Why.  An SS-mode tcgen05.mma (M=128, K=16) costs ~34 + 0.36*N clk (tests/ubench_mma.cu: 39 clk at N=16,
// 44 at N=48, 56 at N=96): it is bound by reading its 4 KB A tile from shared memory, not by the tensor
// pipe, and the cost does not depend on the alignment of the A start address.  The plain one-MMA-per-tap
// formulation (conv_tc.cu) therefore pays 27 x 39 clk per 128 rows whatever Cout is.  Here the three d0
// taps are folded into N:  B = [W(0,k1,k2) | W(1,k1,k2) | W(2,k1,k2)]  (N = 3*CB columns), so one A tile
// feeds three taps -- 9 MMAs per 16 input channels instead of 27:
//     P_k0[x0][p] = sum_{k1,k2,ci} W[k0,k1,k2][co][ci] * X[ci][x0][p + (k1-1)*H2 + (k2-1)]   (tensor core)
//     out[o0][p]  = P_0[o0-1][p] + P_1[o0][p] + P_2[o0+1][p]                                 (epilogue)
// with p a padded-linear position of an (H1 x H2) halo slab of ONE d0 plane.  A CTA sweeps a column of
// the window along d0: plane x0 is one TMA box per channel group into a shared-memory ring, its slab is
// 1-2 GEMM tiles of 128 rows, and the accumulators of three consecutive planes live in a ring of TMEM
// slots.  Because a tile never mixes planes, the d0 shift of the epilogue is a *slot* shift -- the three
// partial sums of an output voxel sit in the same TMEM lane of three slots: no shuffles, no exchange between
// warps, no overlap rows, and no halo planes (planes outside the window are zero: their MMAs are skipped).
// (The first N-fold version folded the d2 taps: the lane shift cost ~900 instructions per warp and tile and
// made the epilogue, not the tensor core, the bottleneck -- profiles/r01_ps_notes.md.)
//
// Persistent, one CTA per SM: 8 or 16 epilogue warps (groups of four = the TMEM lane quarters; a group owns one
// tile of the slab on every 2nd / 4th plane), one TMA producer warp (plane ring, weights resident in shared
// memory), two MMA issuer warps (one elected lane each, alternating planes).  Identity residuals (up-path units, head)
// are read from the plane ring.  All mbarrier waits are bounded (error flag + trap).
#include "conv_tc.cuh"
#include "tc_ptx.cuh"

#include <algorithm>
#include <stdlib.h>
#include <string.h>

namespace sgm {
namespace tc {

namespace {
using namespace tcptx;

constexpr int kIssuers = 2;     // MMA issuer warps (planes are dealt round-robin)
__host__ __device__ constexpr int ps_groups(int cb) { return cb <= 16 ? 4 : 2; }  // epilogue groups (register budget: 104 / 168 per thread)
__host__ __device__ constexpr int ps_threads(int cb) { return (4 * ps_groups(cb) + 1 + kIssuers) * 32; }
constexpr int kPsSmemMax = 227 * 1024;
constexpr int kPadPos = 64;    // positions (16 B each) in front of / behind the plane ring the shifted tiles may touch
constexpr int kRingMax = 16;

struct PsArgs {
  int D[3];                 // tensor extents (input == output)
  int t1, t2, H1, H2, nt1, nt2;
  int H12, PS;              // positions of one plane slab; ring stride per channel group (16-byte units)
  int m;                    // GEMM tiles per plane slab (1 or 2)
  uint32_t mH2;             // ceil(2^32 / H2): exact division of positions < 2^16
  int nunits, units_per_win;
  int R;                    // plane ring depth
  int S_log2, slot_stride;  // TMEM accumulator ring: 2^S_log2 slots of slot_stride columns
  int cgA, c_real, act;      // c_real: real output channels (<= CB)
  float alpha;
  int res_mode;             // 0 none, 1 global CG8 tensor, 2 identity (centre of the plane slab)
  int pl_weighted;
  int rmw;                  // planar output accumulates into the blend buffer: acc += logits * importance (one window)
  int ad0, ad1, ad2, wo[3]; // rmw: accumulator extent and window origin
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

// NC consecutive 32-bit TMEM columns of this lane -> registers (no wait)
template <int NC>
__device__ __forceinline__ void ld_n(uint32_t taddr, uint32_t* v) {
  if constexpr (NC >= 16) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
    ld_n<NC - 16>(taddr + 16, v + 16);
  } else if constexpr (NC >= 8) {
    tc_ld_x8(taddr, v);
    ld_n<NC - 8>(taddr + 8, v + 8);
  } else if constexpr (NC >= 4) {
    tc_ld_x4(taddr, v);
    ld_n<NC - 4>(taddr + 4, v + 4);
  } else if constexpr (NC >= 2) {
    tc_ld_x2(taddr, v);
    ld_n<NC - 2>(taddr + 2, v + 2);
  } else if constexpr (NC == 1) {
    tc_ld_x1(taddr, v);
  }
}

// CB = output channels per d0 tap block (columns k0*CB + co), NCGP = pairs of input channel groups,
// OUTK = OUT_CG8 | OUT_PLANAR
template <int CB, int NCGP, int OUTK>
__global__ void __launch_bounds__(ps_threads(CB), 1)
ps_conv_kernel(const PsArgs a, const __grid_constant__ CUtensorMap tmap) {
  constexpr int NP = (3 * CB + 15) / 16 * 16;  // MMA N
  constexpr int NKB = 9 * NCGP;                // K blocks: (k1, k2) x 16 input channels
  constexpr int CG = 2 * NCGP;                 // input channel groups
  constexpr uint32_t W_BYTES = NKB * NP * 32;
  constexpr int EG = ps_groups(CB);            // epilogue groups of four warps
  constexpr int PW = 4 * EG;                   // producer warp; issuers follow
  constexpr int NTHR = ps_threads(CB);
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  uint8_t* w_smem = smem;
  uint8_t* a_region = smem + W_BYTES;                       // [front pad][ring][tail pad]
  uint8_t* ring = a_region + kPadPos * 16;
  const uint32_t plane_bytes = (uint32_t)CG * a.PS * 16u;   // one ring entry
  const uint32_t a_bytes = (uint32_t)(kPadPos + 128 + kPadPos) * 16u + (uint32_t)a.R * plane_bytes;
  float* bias_s = reinterpret_cast<float*>(a_region + a_bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(bias_s + 32);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 1 + 2 * kRingMax + 32);
  const uint32_t bar0 = smem_u32(bars);
  const uint32_t WBAR = bar0;
  auto PFULL = [&](int s) { return bar0 + 8u * (1 + s); };
  auto PEMPTY = [&](int s) { return bar0 + 8u * (1 + kRingMax + s); };
  auto TFULL = [&](int s) { return bar0 + 8u * (1 + 2 * kRingMax + s); };
  auto TEMPTY = [&](int s) { return bar0 + 8u * (1 + 2 * kRingMax + 16 + s); };
  const int S = 1 << a.S_log2;
  const int D0 = a.D[0];
  const bool tr = a.trace != nullptr && blockIdx.x == 0;

  if (tid == 0) {
    mbar_init(WBAR, 1);
    const uint32_t readers = a.res_mode == 2 ? (a.m == 2 ? 8u : 4u) : 0u;
    for (int s = 0; s < kRingMax; ++s) {
      mbar_init(PFULL(s), 1);
      mbar_init(PEMPTY(s), 1 + readers);  // MMA commit + the epilogue warps that read the identity residual
    }
    for (int s = 0; s < 16; ++s) {
      mbar_init(TFULL(s), 1);
      mbar_init(TEMPTY(s), 12);           // three epilogue steps (hi / mid / lo) x four warps
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (tr) a.trace[0] = clock64();
  }
  if (warp == PW) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // zero the A region once: TMA rewrites positions [0, H12) of every channel group of a ring entry; pads and
  // slack rows (read by shifted tiles, results discarded) stay finite
  for (uint32_t i = tid; i < a_bytes / 16; i += NTHR) reinterpret_cast<uint4*>(a_region)[i] = make_uint4(0, 0, 0, 0);
  if (tid < 32) bias_s[tid] = tid < CB ? __ldg(a.bias + tid) : 0.f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == PW) {
    // ============================ producer: weights once, then plane after plane ============================
    if (lane == 0) {
      mbar_expect_tx(WBAR, W_BYTES);
      bulk_g2s(smem_u32(w_smem), a.w, W_BYTES, WBAR);
      const uint32_t ring_base = smem_u32(ring);
      int pslot = 0;
      uint32_t pphase = 0;
      bool wrapped = false;
      for (int unit = blockIdx.x; unit < a.nunits; unit += gridDim.x) {
        const int n = unit / a.units_per_win;
        const int r = unit - n * a.units_per_win;
        const int b1 = r / a.nt2, b2 = r - b1 * a.nt2;
        for (int x0 = 0; x0 < D0; ++x0) {
          if (wrapped) mbar_wait_or_trap(PEMPTY(pslot), pphase ^ 1u, a.error_flag, 21);
          mbar_expect_tx(PFULL(pslot), (uint32_t)(CG * a.H12 * 16));
          // one 4-D box {H2*8 elements, H1, 1 plane, 1 group}; coordinates outside the window are zero-filled
          // by the hardware == the conv's zero padding (MONAI convolves every window in isolation)
#pragma unroll
          for (int cg = 0; cg < CG; ++cg)
            tma_load_4d(ring_base + (uint32_t)pslot * plane_bytes + (uint32_t)(cg * a.PS) * 16u, &tmap,
                        (b2 * a.t2 - 1) * 8, b1 * a.t1 - 1, x0, n * CG + cg, PFULL(pslot));
          if (++pslot == a.R) pslot = 0, pphase ^= 1u, wrapped = true;
        }
      }
    }
  } else if (warp > PW) {
    // ============================ MMA issuers (the last two warps): one elected lane each ============================
    if (elect_one()) {
      // instruction descriptor: D=f32, A=B=bf16, both K-major, N>>3 at [17,23), M>>4 at [24,29)
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NP >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t ring16 = smem_u32(ring) >> 4;
      const uint32_t w_base16 = smem_u32(w_smem) >> 4;
      const uint32_t d_hi = 8u | (1u << 14);                      // SBO = 128 B: 8 consecutive positions
      const uint32_t a_lbo = ((uint32_t)a.PS & 0x3FFFu) << 16;    // next 8 input channels: next channel group
      const uint32_t b_lbo = ((uint32_t)NP & 0x3FFFu) << 16;      // next 8 input channels of the filter block
      const uint32_t plane16 = (uint32_t)(CG * a.PS);
      mbar_wait_or_trap(WBAR, 0u, a.error_flag, 22);
      // The tcgen05 queue holds only a few MMAs (~200 clk of work), so whatever the issuer does between the last
      // MMA of one tile and the first of the next must be short, or the tensor pipe drains (measured: ~250 idle
      // clk per tile with a naive loop, profiles/r01_ps_notes.md).  Hence (1) two issuer warps alternate planes,
      // (2) the barriers of the NEXT tile are polled in the middle of the current tile's MMA burst, (3) operand
      // offsets are loop-invariant registers.
      const int iw = warp - PW - 1;
      const int my_units = blockIdx.x < a.nunits ? (a.nunits - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
      const int total_planes = my_units * D0;  // planes this CTA sweeps, in producer order
      uint32_t a_off[NKB], b_lo[NKB];
#pragma unroll
      for (int kb = 0; kb < NKB; ++kb) {
        const int k1 = kb / (3 * NCGP), k2 = (kb / NCGP) % 3, cp = kb % NCGP;
        a_off[kb] = (uint32_t)(cp * 2 * a.PS + (k1 - 1) * a.H2 + (k2 - 1));
        b_lo[kb] = ((w_base16 + (uint32_t)(kb * NP * 2)) & 0x3FFFu) | b_lbo;
      }
      int pslot = iw;  // ring entry / phase of plane pc (kIssuers <= ring depth)
      uint32_t pphase = 0;
      bool rdy_p = false, rdy_t = false;  // barriers of the upcoming tile already seen complete
      long long wait_p = 0, wait_t = 0;   // SGM_TRACE: cycles issuer 0 waited for planes / accumulator slots
      int stamps = 0;
      for (int pc = iw; pc < total_planes; pc += kIssuers) {
        // ring entry of this issuer's next plane
        int npslot = pslot + kIssuers;
        uint32_t npphase = pphase;
        if (npslot >= a.R) npslot -= a.R, npphase ^= 1u;
        for (int j = 0; j < a.m; ++j) {
          const int T = pc * a.m + j;  // tile counter of the CTA: accumulator slot T % S, use T / S
          const int tslot = T & (S - 1);
          long long tw0 = 0;
          if (j == 0 && !rdy_p) {
            if (tr) tw0 = clock64();
            mbar_wait_or_trap(PFULL(pslot), pphase, a.error_flag, 23);
            if (tr) wait_p += clock64() - tw0;
          }
          if (!rdy_t && T >= S) {
            if (tr) tw0 = clock64();
            mbar_wait_or_trap(TEMPTY(tslot), ((uint32_t)(T >> a.S_log2) & 1u) ^ 1u, a.error_flag, 24);
            if (tr) wait_t += clock64() - tw0;
          }
          tc_fence_after();
          const uint32_t dcol = tmem_base + (uint32_t)(tslot * a.slot_stride);
          // start addresses stay below 2^14 sixteen-byte units (shared memory < 256 KB): no masking needed
          const uint32_t base = (ring16 + (uint32_t)pslot * plane16 + (uint32_t)(j * 128)) | a_lbo;
          constexpr int kSplit = NKB / 2;
#pragma unroll
          for (int kb = 0; kb < kSplit; ++kb)
            tc_mma(dcol, ((uint64_t)d_hi << 32) | (base + a_off[kb]), ((uint64_t)d_hi << 32) | b_lo[kb], idesc, kb > 0 ? 1u : 0u);
          // poll the barriers of this issuer's next tile while the queued MMAs execute
          if (j + 1 < a.m) {
            const int Tn = T + 1;
            rdy_p = true;
            rdy_t = Tn < S || mbar_try_wait(TEMPTY(Tn & (S - 1)), ((uint32_t)(Tn >> a.S_log2) & 1u) ^ 1u);
          } else if (pc + kIssuers < total_planes) {
            const int Tn = (pc + kIssuers) * a.m;
            rdy_p = mbar_try_wait(PFULL(npslot), npphase);
            rdy_t = Tn < S || mbar_try_wait(TEMPTY(Tn & (S - 1)), ((uint32_t)(Tn >> a.S_log2) & 1u) ^ 1u);
          }
#pragma unroll
          for (int kb = kSplit; kb < NKB; ++kb)
            tc_mma(dcol, ((uint64_t)d_hi << 32) | (base + a_off[kb]), ((uint64_t)d_hi << 32) | b_lo[kb], idesc, 1u);
          tc_commit(TFULL(tslot));
        }
        tc_commit(PEMPTY(pslot));  // this issuer's MMAs that read the plane have completed when this arrives
        pslot = npslot, pphase = npphase;
        if (tr && iw == 0 && stamps < 8) a.trace[1 + stamps++] = clock64();
      }
      if (tr && iw == 0) a.trace[10] = wait_p, a.trace[11] = wait_t, a.trace[12] = clock64();
    }
    __syncwarp();
  } else {
    // ============================ epilogue warps 0..4*EG-1 ============================
    // Group g (four warps = the four TMEM lane quarters) owns tile j of the slab on every pstride-th plane.  One
    // step (an output plane of the tile) is ~200 dependent instructions per warp, so the steps of a tile stream
    // are spread over as many groups as the register file allows.
    const int egroup = warp >> 2, quarter = warp & 3;
    const uint32_t tlane = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const int m = a.m, R = a.R, S_log2 = a.S_log2, slot_stride = a.slot_stride;
    const long long plane_vox = (long long)a.D[1] * a.D[2];
    const long long vox = plane_vox * D0;
    const int j = m == 2 ? (egroup & 1) : 0;
    const int pstride = m == 2 ? EG / 2 : EG, poff = m == 2 ? (egroup >> 1) : egroup;
    const int q = j * 128 + quarter * 32 + lane;  // position inside the plane slab == GEMM row
    const int h1 = (int)__umulhi((uint32_t)q, a.mH2);
    const int h2 = q - h1 * a.H2;
    const bool in_tile = q < a.H12 && h1 >= 1 && h1 <= a.t1 && h2 >= 1 && h2 <= a.t2;
    const int res_mode = a.res_mode, c_real = a.c_real;
    const bool rmw = OUTK == OUT_PLANAR && a.rmw;
    const bool weighted = OUTK == OUT_PLANAR && (a.pl_weighted || rmw);
    const bool act = OUTK == OUT_CG8 && a.act;  // the planar (head) instantiation is conv-only
    const float alpha = a.alpha;
    float bias_r[CB];
#pragma unroll
    for (int c = 0; c < CB; ++c) bias_r[c] = bias_s[c];
    int tbase = 0, pbase = 0;  // tile / plane counters of this CTA at the start of the unit
    bool ok = true;
    for (int unit = blockIdx.x; unit < a.nunits && ok; unit += gridDim.x, tbase += D0 * m, pbase += D0) {
      const int n = unit / a.units_per_win;
      const int r = unit - n * a.units_per_win;
      const int b1 = r / a.nt2, b2 = r - b1 * a.nt2;
      const int r1 = b1 * a.t1 - 1 + h1, r2 = b2 * a.t2 - 1 + h2;
      const bool valid = in_tile && r1 < a.D[1] && r2 < a.D[2];
      // output position = o * oplane + opos12 (rmw: inside the accumulator volume, planes clipped to [0, ad0))
      const long long oplane = rmw ? (long long)a.ad1 * a.ad2 : plane_vox;
      const long long opos12 = rmw ? (long long)(a.wo[1] + r1) * a.ad2 + (a.wo[2] + r2) + (long long)a.wo[0] * oplane
                                   : (long long)r1 * a.D[2] + r2;
      float w1 = 1.f, w2 = 1.f;
      if (weighted && valid) w1 = __ldg(a.imap1 + r1), w2 = __ldg(a.imap2 + r2);
      int ps = (pbase + poff) % R;  // ring entry of plane o and the phase of its "landed" barrier
      uint32_t pph = (uint32_t)((pbase + poff) / R) & 1u;
      for (int o = poff; o < D0; o += pstride) {
        const int t_mid = tbase + o * m + j, t_lo = t_mid - m, t_hi = t_mid + m;
        const bool has_lo = o > 0, has_hi = o + 1 < D0;
        const long long opos = (long long)o * oplane + opos12;
        uint4 gres[(CB + 7) / 8];
        if (res_mode == 1) {
#pragma unroll
          for (int pc = 0; pc < (CB + 7) / 8; ++pc) {
            gres[pc] = make_uint4(0, 0, 0, 0);
            if (valid) gres[pc] = __ldg(reinterpret_cast<const uint4*>(a.res + (((long long)n * a.cgA + pc) * vox + opos) * 8));
          }
        }
        float imw = 1.f;
        if (weighted) imw = fmaxf(__fmul_rn(__fmul_rn(__ldg(a.imap0 + o), w1), w2), a.imap_floor);
        // planes o-1, o, o+1 belong to different issuers and this group skips planes: observe the barrier of every
        // tile read (only the newest one ever blocks)
        if (has_hi) ok = mbar_wait(TFULL(t_hi & (S - 1)), (uint32_t)(t_hi >> S_log2) & 1u, a.error_flag, 25);
        if (ok) ok = mbar_wait(TFULL(t_mid & (S - 1)), (uint32_t)(t_mid >> S_log2) & 1u, a.error_flag, 26);
        if (ok && has_lo) ok = mbar_wait(TFULL(t_lo & (S - 1)), (uint32_t)(t_lo >> S_log2) & 1u, a.error_flag, 27);
        if (!ok) break;
        tc_fence_after();
        // ---- the three partial sums of this lane's voxel: same lane, three accumulator slots
        const int s_lo = t_lo & (S - 1), s_mid = t_mid & (S - 1), s_hi = t_hi & (S - 1);
        uint32_t lo[CB], mid[CB], hi[CB];
        // NOTE: every tcgen05.ld is followed by its tcgen05.wait::ld inside the SAME basic block.  The compiler
        // treats the asm outputs as ready immediately; a register move inserted at a control-flow merge between
        // the load and the wait would read the destination before the data has landed.
        if (has_lo && has_hi) {
          ld_n<CB>(tlane + (uint32_t)(s_lo * slot_stride), lo);
          ld_n<CB>(tlane + (uint32_t)(s_mid * slot_stride) + CB, mid);
          ld_n<CB>(tlane + (uint32_t)(s_hi * slot_stride) + 2 * CB, hi);
          tc_ld_wait();
        } else {  // first / last plane of the column: the missing neighbour plane is zero
#pragma unroll
          for (int c = 0; c < CB; ++c) lo[c] = 0u, hi[c] = 0u;
          ld_n<CB>(tlane + (uint32_t)(s_mid * slot_stride) + CB, mid);
          tc_ld_wait();
          if (has_lo) {
            ld_n<CB>(tlane + (uint32_t)(s_lo * slot_stride), lo);
            tc_ld_wait();
          }
          if (has_hi) {
            ld_n<CB>(tlane + (uint32_t)(s_hi * slot_stride) + 2 * CB, hi);
            tc_ld_wait();
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {  // every slot collects 3 steps x 4 warps; the first / last plane of a column stand in
          mbar_arrive(TEMPTY(s_mid));                      // for the step that does not exist
          mbar_arrive(TEMPTY(has_lo ? s_lo : s_mid));
          mbar_arrive(TEMPTY(has_hi ? s_hi : s_mid));
        }
        constexpr int CBP = (CB + 7) / 8 * 8;  // CG8 tensors carry whole channel groups (padding channels: conv part 0)
        float v[CBP];
#pragma unroll
        for (int c = 0; c < CBP; ++c)
          v[c] = c < CB ? (__uint_as_float(lo[c]) + __uint_as_float(mid[c])) + __uint_as_float(hi[c]) + bias_r[c] : 0.f;
        if (act) {
#pragma unroll
          for (int c = 0; c < CB; ++c) v[c] = prelu(v[c], alpha);
        }
        if (res_mode == 1) {
#pragma unroll
          for (int pc = 0; pc < (CB + 7) / 8; ++pc) {
            float rr[8];
            unpack8(gres[pc], rr);
#pragma unroll
            for (int c = 0; c < 8; ++c) v[8 * pc + c] += rr[c];
          }
        } else if (res_mode == 2) {
          // generic-proxy reads of TMA-written data: observe the plane's own barrier (it completed long ago)
          ok = mbar_wait(PFULL(ps), pph, a.error_flag, 28);
          if (!ok) break;
          const uint8_t* pl = ring + (size_t)ps * plane_bytes + (size_t)q * 16;
#pragma unroll
          for (int pc = 0; pc < (CB + 7) / 8; ++pc) {
            float rr[8];
            unpack8(*reinterpret_cast<const uint4*>(pl + (size_t)pc * a.PS * 16), rr);
#pragma unroll
            for (int c = 0; c < 8; ++c) v[8 * pc + c] += rr[c];
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(PEMPTY(ps));  // this warp no longer reads the plane
        }
        if (valid && !(rmw && (a.wo[0] + o < 0 || a.wo[0] + o >= a.ad0))) {
          if (OUTK == OUT_CG8) {
            __nv_bfloat16* dst = a.out + ((long long)n * a.cgA * vox + opos) * 8;
#pragma unroll
            for (int g = 0; g < (CB + 7) / 8; ++g) {
              float o8[8];
#pragma unroll
              for (int c = 0; c < 8; ++c) o8[c] = v[8 * g + c];
              if (g < a.cgA) *reinterpret_cast<uint4*>(dst + (long long)g * vox * 8) = pack8(o8);
            }
            for (int g = (CB + 7) / 8; g < a.cgA; ++g)  // channel padding of the CG8 tensor stays zero
              *reinterpret_cast<uint4*>(dst + (long long)g * vox * 8) = make_uint4(0, 0, 0, 0);
          } else {
            float* dst = a.pl_out + (long long)n * a.pl_nstride + opos;
            const long long cs = a.pl_cstride;
            if (!rmw) {
#pragma unroll
              for (int c = 0; c < CB; ++c, dst += cs)
                if (c < c_real) __stcs(dst, weighted ? __fmul_rn(v[c], imw) : v[c]);
            } else {  // seg *= w; out += seg (two roundings, as MONAI); windows are launched one after another
              float oldv[CB];
#pragma unroll
              for (int c = 0; c < CB; ++c)
                if (c < c_real) oldv[c] = __ldcg(dst + c * cs);
#pragma unroll
              for (int c = 0; c < CB; ++c)
                if (c < c_real) __stcg(dst + c * cs, __fadd_rn(oldv[c], __fmul_rn(v[c], imw)));
            }
          }
        }
        ps += pstride;
        if (ps >= R) ps -= R, pph ^= 1u;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == PW) {
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

struct PsPlan {
  int key[5];
  PsArgs args;
  int smem_bytes, grid;
};

int ps_np(int cb) { return round_up(3 * cb, 16); }

// column block per d0 tap for a conv with `cout` real output channels (one template instantiation each)
int ps_cb(int cout) {
  static const int cbs[] = {4, 8, 10, 16, 20, 24, 32};
  for (int cb : cbs)
    if (cout <= cb) return cb;
  return 0;
}

int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    // multi-GPU runs leave a few SMs to the NCCL point-to-point kernels of the seam exchange: a persistent grid of
    // one CTA per SM would otherwise need a second wave whenever NCCL holds an SM
    if (const char* env = getenv("SGM_RESERVE_SMS")) n = std::max(8, n - atoi(env));
  }
  return n;
}

int ps_fixed_smem(int cb, int ncgp) {
  const int w_bytes = 9 * ncgp * ps_np(cb) * 32;
  return w_bytes + (kPadPos + 128 + kPadPos) * 16 + 32 * 4 + (1 + 2 * kRingMax + 32) * 8 + 16 + 128;
}

int ps_plan(const TcConv& c, const TcIO& io, PsPlan& pl) {
  PsArgs& a = pl.args;
  memset(&a, 0, sizeof(a));
  const int CG = 2 * c.ps_ncgp, NP = ps_np(c.ps_cb);
  for (int i = 0; i < 3; ++i) a.D[i] = io.od[i];
  int stride = 32;
  while (stride < NP) stride <<= 1;
  a.slot_stride = stride;
  const int S = 512 / stride;  // 16, 8 or 4 accumulator slots
  a.S_log2 = S == 16 ? 4 : (S == 8 ? 3 : 2);
  const int m_max = std::min(2, (S - 2) / 2);  // three planes of m tiles live + MMA run-ahead
  const int nsm = sm_count();
  const double t_tile = 9.0 * c.ps_ncgp * (34.0 + 0.36 * NP);
  double best = 1e30;
  int bt1 = 0, bt2 = 0, bm = 0;
  for (int t1 = 1; t1 <= std::min(a.D[1], 62); ++t1)
    for (int t2 = 1; t2 <= std::min(a.D[2], 30); ++t2) {  // TMA box: H2 * 8 elements <= 256
      const int H1 = t1 + 2, H2 = t2 + 2;
      const int m = ceil_div(H1 * H2, 128);
      if (m > m_max) continue;
      const long long units = (long long)ceil_div(a.D[1], t1) * ceil_div(a.D[2], t2) * io.n;
      const double waves = (double)((units + nsm - 1) / nsm);
      // per unit: D0 planes of m tiles + pipeline fill; small bias towards long d2 runs (coalescing)
      const double cost = waves * ((double)a.D[0] * m * t_tile + 3000.0) * (1.0 + 0.02 / t2);
      if (cost < best) best = cost, bt1 = t1, bt2 = t2, bm = m;
    }
  SGM_REQUIRE(bt1 > 0, SGM_ERR_UNSUPPORTED, "ps_plan: no slab shape fits");
  a.t1 = bt1, a.t2 = bt2, a.H1 = bt1 + 2, a.H2 = bt2 + 2, a.m = bm;
  a.nt1 = ceil_div(a.D[1], bt1), a.nt2 = ceil_div(a.D[2], bt2);
  a.H12 = a.H1 * a.H2;
  a.PS = round_up(a.H12, 8);
  a.mH2 = (uint32_t)((0x100000000ULL + a.H2 - 1) / a.H2);
  a.units_per_win = a.nt1 * a.nt2;
  a.nunits = a.units_per_win * io.n;
  const int fixed = ps_fixed_smem(c.ps_cb, c.ps_ncgp);
  const int plane_bytes = CG * a.PS * 16;
  a.R = std::min(kRingMax, (kPsSmemMax - fixed) / plane_bytes);
  SGM_REQUIRE(a.R >= 4, SGM_ERR_UNSUPPORTED, "ps_plan: plane ring does not fit shared memory");
  a.R = std::min(a.R, 12);
  pl.smem_bytes = fixed + a.R * plane_bytes;
  pl.grid = std::min(nsm, a.nunits);
  return SGM_OK;
}

template <int CB, int NCGP, int OUTK>
int launch_t(const PsArgs& a, const CUtensorMap& tm, int grid, int smem, cudaStream_t st) {
  static DeviceOnce attr_once;
  if (attr_once.first()) {
    SGM_CUDA_CHECK(cudaFuncSetAttribute(ps_conv_kernel<CB, NCGP, OUTK>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPsSmemMax));
  }
  ps_conv_kernel<CB, NCGP, OUTK><<<grid, ps_threads(CB), smem, st>>>(a, tm);
  SGM_CUDA_CHECK(cudaGetLastError());
  return SGM_OK;
}

template <int OUTK>
int launch_cb(int cb, int ncgp, const PsArgs& a, const CUtensorMap& tm, int grid, int smem, cudaStream_t st) {
#define SGM_PS_CASE(CBV, NCGPV) \
  if (cb == CBV && ncgp == NCGPV) return launch_t<CBV, NCGPV, OUTK>(a, tm, grid, smem, st)
  SGM_PS_CASE(4, 1);
  SGM_PS_CASE(8, 1);
  SGM_PS_CASE(10, 1);
  SGM_PS_CASE(16, 1);
  SGM_PS_CASE(20, 2);
  SGM_PS_CASE(24, 2);
  SGM_PS_CASE(32, 2);
#undef SGM_PS_CASE
  set_error("ps_launch: no instantiation for %d output channels per tap / %d channel pairs", cb, ncgp);
  return SGM_ERR_UNSUPPORTED;
}

}  // namespace

int ps_pack(const sgm_conv_desc& d, TcConv* c) {
  c->ps_cb = 0;
  if (getenv("SGM_NO_PS")) return SGM_OK;
  if (d.kind != SGM_KIND_CONV || d.kernel != 3 || d.stride != 1 || c->flat0 || c->mode != MODE_S1) return SGM_OK;
  if (d.cout > 32 || d.cin > 32) return SGM_OK;
  const int cb = ps_cb(d.cout);
  const int ncgp = c->cgin / 2;
  if (!cb || (cb <= 16) != (ncgp == 1) || ncgp > 2) return SGM_OK;  // instantiated: <=16 ch x 1 pair, 17..32 x 2 pairs
  const int NP = ps_np(cb), NKB = 9 * ncgp;
  std::vector<uint16_t> w((size_t)NKB * 2 * NP * 8, 0);
  for (int k1 = 0; k1 < 3; ++k1)
    for (int k2 = 0; k2 < 3; ++k2)
      for (int cp = 0; cp < ncgp; ++cp) {
        const int kb = (k1 * 3 + k2) * ncgp + cp;
        for (int kc = 0; kc < 2; ++kc)
          for (int k0 = 0; k0 < 3; ++k0)
            for (int co = 0; co < d.cout; ++co)
              for (int k8 = 0; k8 < 8; ++k8) {
                const int ci = (cp * 2 + kc) * 8 + k8;
                if (ci >= d.cin) continue;
                const int tap = (k0 * 3 + k1) * 3 + k2;
                w[(((size_t)kb * 2 + kc) * NP + (k0 * cb + co)) * 8 + k8] =
                    f2bf(d.weight[((size_t)co * d.cin + ci) * 27 + tap]);
              }
      }
  if (cudaMalloc(&c->ps_w, w.size() * 2) != cudaSuccess) {
    set_error("ps_pack: cudaMalloc failed: %s", cudaGetErrorString(cudaGetLastError()));
    return SGM_ERR_CUDA;
  }
  SGM_CUDA_CHECK(cudaMemcpy(c->ps_w, w.data(), w.size() * 2, cudaMemcpyHostToDevice));
  c->ps_cb = cb, c->ps_ncgp = ncgp;
  c->ps_plan_cache = new std::vector<PsPlan>();
  return SGM_OK;
}

void ps_free(TcConv* c) {
  if (c->ps_w) cudaFree(c->ps_w);
  c->ps_w = nullptr;
  delete reinterpret_cast<std::vector<PsPlan>*>(c->ps_plan_cache);
  c->ps_plan_cache = nullptr;
}

bool ps_applicable(const TcConv& c, const TcIO& io) {
  if (!c.ps_cb || !c.ps_w || !tma_available()) return false;
  if (io.in1 || io.cg1 || io.outB) return false;
  if (io.out_kind == OUT_BLEND && io.n != 1) return false;  // read-modify-write blend: one window per launch
  if (io.out_kind != OUT_CG8 && c.actA) return false;  // the planar instantiation is conv-only (MONAI's head)
  if (io.cg0 != 2 * c.ps_ncgp) return false;
  for (int i = 0; i < 3; ++i)
    if (io.id[i] != io.od[i]) return false;
  if ((long long)io.od[0] * io.od[1] * io.od[2] >= (1LL << 31)) return false;
  return true;
}

int ps_launch(const TcConv& c, const TcIO& io, int* error_flag_dev, cudaStream_t st) {
  auto* plans = reinterpret_cast<std::vector<PsPlan>*>(c.ps_plan_cache);
  const int key[5] = {io.od[0], io.od[1], io.od[2], io.n, 0};
  const PsPlan* pe = nullptr;
  for (auto& e : *plans)
    if (memcmp(e.key, key, sizeof(key)) == 0) pe = &e;
  if (!pe) {
    PsPlan e;
    memcpy(e.key, key, sizeof(key));
    int rc = ps_plan(c, io, e);
    if (rc) return rc;
    plans->push_back(e);
    pe = &plans->back();
  }
  PsArgs a = pe->args;
  a.cgA = io.cgA, a.c_real = c.c_real, a.act = c.actA, a.alpha = c.alphaA;
  a.w = c.ps_w, a.bias = c.bias;
  a.out = (__nv_bfloat16*)io.outA, a.res = (const __nv_bfloat16*)io.res;
  a.res_mode = 0;
  if (io.res) a.res_mode = (io.res == io.in0 && io.cgA == 2 * c.ps_ncgp) ? 2 : 1;
  a.pl_weighted = io.pl_weighted;
  a.rmw = io.out_kind == OUT_BLEND;
  a.ad0 = io.ad0, a.ad1 = io.ad1, a.ad2 = io.ad2;
  for (int i = 0; i < 3; ++i) a.wo[i] = io.wo[i];
  a.pl_out = io.pl_out, a.pl_cstride = io.pl_cstride, a.pl_nstride = a.rmw ? 0 : io.pl_nstride;
  a.imap0 = io.imap[0], a.imap1 = io.imap[1], a.imap2 = io.imap[2], a.imap_floor = io.imap_floor;
  a.error_flag = error_flag_dev;
  static const bool dbg = getenv("SGM_DEBUG") != nullptr;
  if (dbg)
    fprintf(stderr,
            "[ps_launch] CB=%d NCGP=%d D=(%d,%d,%d) n=%d t=(%d,%d) H12=%d PS=%d m=%d units=%d grid=%d smem=%d R=%d "
            "slots=%d x %d res=%d out=%d\n",
            c.ps_cb, c.ps_ncgp, a.D[0], a.D[1], a.D[2], io.n, a.t1, a.t2, a.H12, a.PS, a.m, a.nunits, pe->grid,
            pe->smem_bytes, a.R, 1 << a.S_log2, a.slot_stride, a.res_mode, io.out_kind);
  static const bool trace_on = getenv("SGM_TRACE") != nullptr;
  static long long* trace_dev = nullptr;
  if (trace_on) {
    if (!trace_dev) cudaMalloc(&trace_dev, 32 * sizeof(long long));
    cudaMemsetAsync(trace_dev, 0, 32 * sizeof(long long), st);
    a.trace = trace_dev;
  }
  CUtensorMap tm;
  const int box[3] = {1, a.H1, a.H2};
  int rc = make_brick_map(&tm, io.in0, io.n * io.cg0, io.id, box);
  if (rc) return rc;
  if (io.out_kind == OUT_CG8) rc = launch_cb<OUT_CG8>(c.ps_cb, c.ps_ncgp, a, tm, pe->grid, pe->smem_bytes, st);
  else rc = launch_cb<OUT_PLANAR>(c.ps_cb, c.ps_ncgp, a, tm, pe->grid, pe->smem_bytes, st);
  if (rc) return rc;
  if (trace_on) {
    long long t[32];
    cudaStreamSynchronize(st);
    cudaMemcpy(t, trace_dev, sizeof(t), cudaMemcpyDeviceToHost);
    auto d = [&](int i) { return t[i] ? (double)(t[i] - t[0]) : -1.0; };
    fprintf(stderr,
            "[ps trace] CB=%d D=(%d,%d,%d) n=%d t=(%d,%d) m=%d units/cta=%.1f | plane issued %.0f %.0f %.0f %.0f %.0f %.0f "
            "%.0f %.0f | end %.0f cycles; issuer waited %.0f for planes, %.0f for TMEM slots, done at %.0f\n",
            c.ps_cb, a.D[0], a.D[1], a.D[2], io.n, a.t1, a.t2, a.m, (double)a.nunits / pe->grid, d(1), d(2), d(3), d(4),
            d(5), d(6), d(7), d(8), d(9), (double)t[10], (double)t[11], d(12));

  }
  return SGM_OK;
}

}  // namespace tc
}  // namespace sgm
