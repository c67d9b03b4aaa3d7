// "Row-sweep" tcgen05 implicit-GEMM convolution for the head: stride-1 3x3x3 conv with <= 16 input and <= 10 output
// channels at full window resolution (C -> C at 96^3: 29 % of the UNet's FLOPs), planar fp32 output (logits or
// importance-weighted logits for the deferred blend), optional identity residual.
//
// Why.  An SS-mode tcgen05.mma (M = 128, K = 16) costs ~34 + 0.36 N clk (profiles/r01_ps_notes.md): it is bound by
// reading its A tile, so the plane-sweep kernel (conv_ps.cu: d0 taps folded into N, 9 MMAs of N = 32 per 128 rows)
// spends ~360 clk per tile for 30 useful columns.  Here BOTH the d0 and the d1 taps are folded into N -- 3 MMAs of
// N = 96 (one per d2 tap) per 128 positions, 168 clk -- and, unlike an epilogue-side fold, the partial sums never
// have to be combined by the epilogue warps: a GEMM tile is ONE input row (x0, r) of the window (its D2 <= 128
// positions are the 128 TMEM lanes), and the accumulator columns are laid out as
//     column(out row o, plane slot p, channel c) = o * RP + p * CP + c,   RP = 3*CP rounded up to 16,  p = (virtual out plane) mod 3,
// so the nine (k0, k1) blocks of B = [k1=2: p0 p1 p2 | k1=1: p0 p1 p2 | k1=0: p0 p1 p2] land, with ONE accumulating MMA
// whose D address is column(r - 1, 0, 0), on the three output rows r-1, r, r+1 of the three output planes x0-1, x0,
// x0+1.  Consecutive input rows write OVERLAPPING column ranges; the tensor pipe executes MMAs in issue order, so the
// overlap is an ordinary accumulate chain.  Which k0 sits in which slot depends on x0 mod 3: the packed weights hold
// the three rotations (3 x 9 KB).  Each k1 block is padded to RP rows with zero weights, so N = 3*RP is a multiple of 16
// and the padding columns of a row only ever receive +0.0 (nobody else touches them: no read-modify-write race with the
// epilogue's clears); the first / last two rows of a strip use the trailing / leading RP or 2*RP rows of the same B tile
// (the start address of a K-major SWIZZLE_NONE operand may be any multiple of 16 B).
//
// A persistent CTA (one per SM) sweeps a strip of t1 output rows of one window along d0.  An output plane is complete
// once the next input plane has been multiplied; its epilogue (16 warps = 4 groups x the 4 TMEM lane quarters, rows
// dealt round-robin to the groups) reads CP columns per row, ZEROES them (tcgen05.st) for the plane that reuses the
// slot three planes later, and signals a per-row "cleared" barrier the MMA issuer observes before it touches the row
// again -- row granularity, so the MMAs of plane x0+1 run while plane x0-1 drains.  In the other direction the issuer
// commits "multiplied" barriers per chunk of four input rows (tcgen05.commit costs ~400 clk of issue time).  Every (window, strip) is framed
// by two phantom output planes (-1 and D0) that only clear the garbage the first / last input plane deposits outside
// the window, and by two phantom input planes (D0, D0+1) that carry no MMAs but still commit the per-row barriers: no
// special-cased weights, and every barrier has exactly one phase per virtual plane, observed in order by its waiter
// (mbarrier parity waits alias when a waiter falls two phases behind -- the first version let the epilogue run through
// the last planes of a unit ungated and deadlocked; tests/sim_rs_protocol.py is the randomised model of the protocol).
// All mbarrier waits are bounded and fail soft (error code in the flag sgm_unet_check reads, CTA-wide abort).
#include "conv_tc.cuh"
#include "tc_ptx.cuh"

#include <algorithm>
#include <stdlib.h>
#include <string.h>

namespace sgm {
namespace tc {

namespace {
using namespace tcptx;

constexpr int kRsEpiWarps = 16;
constexpr int kRsIssuers = 2;                        // MMA issuer warps (alternating input rows)
constexpr int kRsThreads = (kRsEpiWarps + 1 + kRsIssuers) * 32;  // + producer warp + issuer warps
constexpr int kRsSmemMax = 227 * 1024;
constexpr int kRsRingMax = 6;
constexpr int kRsRowsMax = 16;                       // output rows per strip (TMEM: t1 * RP <= 512 columns)
constexpr int kRsTailPad = 2304;                     // bytes behind the ring a shifted 128-row tile may touch

struct RsArgs {
  int D[3];
  int t1, H1, W;       // output rows per strip, input rows (t1 + 2), positions per input row (D2 + 2)
  int GS;              // channel-group slab of a ring entry in 16-byte units
  int nstrips, nunits;
  int R;               // plane ring depth
  int nq;              // TMEM lane quarters that hold real positions: ceil(D2 / 32)
  int issuers;         // MMA issuer warps in use: 1 (deterministic accumulation order, default) or 2 (SGM_RS_ISSUERS=2)
  int c_real;
  int res_mode;        // 0 none, 2 identity (centre of the input plane, read from the ring)
  int pl_weighted;
  const __nv_bfloat16* w;
  const float* bias;
  float* pl_out;
  long long pl_cstride, pl_nstride;
  const float* imap0;
  const float* imap1;
  const float* imap2;
  float imap_floor;
  int* error_flag;
  long long* trace;    // SGM_TRACE: clock stamps / wait cycles of CTA 0
  int abl;             // SGM_RS_ABL: timing ablations of the issuer (wrong results): 8 narrow MMAs, 256 no MMAs, 1 one MMA per row
};

template <int NC>
__device__ __forceinline__ void rs_ld(uint32_t taddr, uint32_t* v) {
  if constexpr (NC >= 8) {
    tc_ld_x8(taddr, v);
    rs_ld<NC - 8>(taddr + 8, v + 8);
  } else if constexpr (NC >= 4) {
    tc_ld_x4(taddr, v);
    rs_ld<NC - 4>(taddr + 4, v + 4);
  } else if constexpr (NC >= 2) {
    tc_ld_x2(taddr, v);
    rs_ld<NC - 2>(taddr + 2, v + 2);
  } else if constexpr (NC == 1) {
    tc_ld_x1(taddr, v);
  }
}

// NC consecutive 32-bit TMEM columns of this warp's lane quarter <- 0 (no wait)
template <int NC>
__device__ __forceinline__ void rs_st_zero(uint32_t taddr) {
  const uint32_t z = 0u;
  if constexpr (NC >= 8) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr), "r"(z) : "memory");
    rs_st_zero<NC - 8>(taddr + 8);
  } else if constexpr (NC >= 4) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %1, %1, %1};" ::"r"(taddr), "r"(z) : "memory");
    rs_st_zero<NC - 4>(taddr + 4);
  } else if constexpr (NC >= 2) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %1};" ::"r"(taddr), "r"(z) : "memory");
    rs_st_zero<NC - 2>(taddr + 2);
  } else if constexpr (NC == 1) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(z) : "memory");
  }
}
__device__ __forceinline__ void rs_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// Bounded wait that fails soft: on a timeout (or when another role already gave up) the first wait code is kept in the
// error flag, the CTA-wide abort flag is raised and the caller unwinds to the final barrier -- the kernel ends, and
// sgm_unet_check reports the code instead of a dead context.
constexpr long long kRsWaitCycles = 2000000000LL;  // ~1 s: far beyond any legitimate wait, short enough to report instead of hanging
__device__ __forceinline__ bool rs_wait(uint32_t bar, uint32_t parity, int* err, int code, volatile int* abort_s,
                                        unsigned sleep_ns = 0) {
  if (mbar_try_wait(bar, parity)) return true;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (sleep_ns) __nanosleep(sleep_ns);  // epilogue warps: leave the issue slots to the MMA issuer while they wait
    if (*abort_s || clock64() - t0 > kRsWaitCycles) {
      atomicCAS(err, 0, code);
      *abort_s = 1;
      return false;
    }
  }
  return true;
}

__host__ __device__ constexpr int rs_pad16(int v) { return (v + 15) / 16 * 16; }
__host__ __device__ constexpr int rs_rp(int cp) { return rs_pad16(3 * cp); }                 // accumulator columns per output row (3 plane slots + padding)
__host__ __device__ constexpr int rs_nb(int cp) { return 3 * rs_rp(cp); }                    // B rows per K chunk: three k1 blocks of RP rows
__host__ __device__ constexpr int rs_w_bytes(int cp) { return 3 * 3 * 2 * rs_nb(cp) * 16; }  // [rot][k2][kchunk][NB][8 bf16]

// CP = accumulator columns per (output row, plane slot) >= CR = real output channels (stores unrolled without guards)
template <int CP, int CR>
__global__ void __launch_bounds__(kRsThreads, 1) rs_conv_kernel(const RsArgs a, const __grid_constant__ CUtensorMap tmap) {
  constexpr int RP = rs_rp(CP);  // columns per output row; the RP - 3*CP padding columns only ever receive +0.0
  constexpr int N3 = RP, N6 = 2 * RP, N9 = 3 * RP;
  constexpr int NB = rs_nb(CP);
  constexpr uint32_t W_BYTES = rs_w_bytes(CP);
  constexpr uint32_t W_REGION = (W_BYTES + 127u) / 128u * 128u;  // TMA destinations are 128-byte aligned
  constexpr uint32_t WROT16 = 3 * 2 * NB;  // one rotation, in 16-byte units
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  uint8_t* w_smem = smem;
  uint8_t* ring = smem + W_REGION;
  const uint32_t plane_bytes = 2u * (uint32_t)a.GS * 16u;  // one ring entry: two channel groups
  const uint32_t a_bytes = (uint32_t)a.R * plane_bytes + kRsTailPad;
  float* bias_s = reinterpret_cast<float*>(ring + a_bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(bias_s + 32);
  const uint32_t bar0 = smem_u32(bars);
  const uint32_t WBAR = bar0;
  auto PFULL = [&](int s) { return bar0 + 8u * (1 + s); };
  auto PEMPTY = [&](int s) { return bar0 + 8u * (1 + kRsRingMax + s); };
  auto FULL = [&](int c) { return bar0 + 8u * (1 + 2 * kRsRingMax + c); };                      // input rows of chunk c multiplied
  auto CLR = [&](int c) { return bar0 + 8u * (1 + 2 * kRsRingMax + (kRsRowsMax + 2) + c); };    // output rows 4c .. 4c+3 cleared
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 1 + 2 * kRsRingMax + (kRsRowsMax + 2) + kRsRowsMax);
  volatile int* abort_s = reinterpret_cast<volatile int*>(tmem_slot + 1);
  const int D0 = a.D[0], t1 = a.t1, H1 = a.H1, R = a.R, nq = a.nq;
  const bool tr = a.trace != nullptr && blockIdx.x == 0;
  constexpr int PW = kRsEpiWarps, IW = kRsEpiWarps + 1;  // producer warp, first issuer warp

  if (tid == 0) {
    *abort_s = 0;
    mbar_init(WBAR, 1);
    const uint32_t readers = a.res_mode == 2 ? (uint32_t)(4 * nq) : 0u;
    for (int s = 0; s < kRsRingMax; ++s) {
      mbar_init(PFULL(s), 1);
      mbar_init(PEMPTY(s), (uint32_t)a.issuers + readers);  // MMA commits + the epilogue warps that read the identity residual
    }
    for (int i = 0; i < kRsRowsMax + 2; ++i) mbar_init(FULL(i), (uint32_t)a.issuers);
    for (int c = 0; c < kRsRowsMax / 4; ++c) mbar_init(CLR(c), (uint32_t)(nq * max(min(4, t1 - 4 * c), 1)));  // every row of the chunk, every lane quarter
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (tr) a.trace[0] = clock64();
  }
  if (warp == PW) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // zero the ring once: slack behind the rows (read by the 128-row tiles, results discarded) stays finite
  for (uint32_t i = tid; i < a_bytes / 16; i += kRsThreads) reinterpret_cast<uint4*>(ring)[i] = make_uint4(0, 0, 0, 0);
  if (tid < 32) bias_s[tid] = tid < CR ? __ldg(a.bias + tid) : 0.f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // the accumulators are an accumulate-only ring: start from zero (each epilogue group clears 128 columns)
  if (warp < kRsEpiWarps) {
    const uint32_t t0 = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 128);
#pragma unroll
    for (int c = 0; c < 128; c += 8) rs_st_zero<8>(t0 + c);
    rs_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  const int my_units = (int)blockIdx.x < a.nunits ? (a.nunits - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (warp == PW) {
    // ============================ producer: weights once, then input plane after input plane ============================
    if (lane == 0) {
      mbar_expect_tx(WBAR, W_BYTES);
      bulk_g2s(smem_u32(w_smem), a.w, W_BYTES, WBAR);
      const uint32_t ring_base = smem_u32(ring);
      const uint32_t tx = 2u * (uint32_t)(H1 * a.W) * 16u;
      int slot = 0;
      uint32_t ephase = 1;  // parity that reads "never used" on a fresh barrier
      bool ok = true;
      for (int kl = 0; kl < my_units && ok; ++kl) {
        const int unit = (int)blockIdx.x + kl * (int)gridDim.x;
        const int n = unit / a.nstrips, strip = unit - n * a.nstrips;
        for (int x0 = 0; x0 < D0; ++x0) {
          ok = rs_wait(PEMPTY(slot), ephase, a.error_flag, 41, abort_s);
          if (!ok) break;
          mbar_expect_tx(PFULL(slot), tx);
          // one 5-D box {8 ch, W positions, H1 rows, 1 plane, 1 group}; coordinates outside the window are zero-filled
          // by the hardware == the conv's zero padding (MONAI convolves every window in isolation)
#pragma unroll
          for (int cg = 0; cg < 2; ++cg)
            tma_load_5d(ring_base + (uint32_t)slot * plane_bytes + (uint32_t)(cg * a.GS) * 16u, &tmap, 0, -1,
                        strip * t1 - 1, x0, n * 2 + cg, PFULL(slot));
          if (++slot == R) slot = 0, ephase ^= 1u;
        }
      }
    }
  } else if (warp >= IW && warp < IW + a.issuers) {
    // ============================ MMA issuer(s): one elected lane per warp ============================
    // Default: ONE issuer -- every output row then receives its 27 x (channel pair) partial products in a fixed order
    // (planes, rows, d2 taps ascending) that does not depend on the strip height, the batch or the timing, so logits
    // are reproducible bit for bit between runs, batch sizes and the multi-GPU partitions.  SGM_RS_ISSUERS=2 shares
    // the rows between two warps (alternating rows): faster issue, but the two instruction streams interleave in the
    // tensor pipe in a timing-dependent order, i.e. the fp32 summation order is no longer fixed.
    // This loop runs on ONE thread next to a tensor pipe that wants a new MMA every ~56 clk, at ~6 clk per dependent
    // instruction: no divisions (the first version spent 500-700 clk per row on `%` and `/`), ring / rotation / parity
    // state carried incrementally, operand descriptors advanced by additions -- and the rows are shared by two issuer
    // warps.  Accumulation commutes and the tensor pipe executes whole MMAs one after another, so the interleaving of
    // the two instruction streams does not matter; each issuer commits its own MMAs ("multiplied" barriers count two).
    if (elect_one()) {
      const int iw = warp - IW, ni = a.issuers;
      auto idesc = [](uint32_t n) { return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((128u >> 4) << 24); };
      const uint32_t id3 = idesc(N3), id6 = idesc(a.abl & 8 ? N3 : N6), id9 = idesc(a.abl & 8 ? N3 : N9);
      const bool no_mma = a.abl & 256, one_mma = a.abl & 1;
      const uint32_t ring16 = smem_u32(ring) >> 4;
      const uint32_t w16 = smem_u32(w_smem) >> 4;
      const uint64_t d_hi = (uint64_t)(8u | (1u << 14)) << 32;  // SBO = 128 B: 8 consecutive positions / B rows
      const uint32_t a_lbo = ((uint32_t)a.GS & 0x3FFFu) << 16;  // next 8 input channels: next channel group
      const uint32_t b_lbo = ((uint32_t)NB & 0x3FFFu) << 16;
      const uint32_t plane16 = 2u * (uint32_t)a.GS;
      const uint32_t W = (uint32_t)a.W;
      bool ok = rs_wait(WBAR, 0u, a.error_flag, 42, abort_s);
      long long wait_p = 0, wait_c = 0;
      if (tr && iw == 0) a.trace[1] = clock64();
      int slot = 0;          // ring entry of the next real input plane
      uint32_t pphase = 0;   // parity of its "landed" barrier
      uint32_t rot = 0;      // virtual output plane index of (xi - 1), mod 3: which weight rotation
      uint32_t vpar = 1;     // parity of the "cleared" phase to observe (virtual plane counter of the CTA - 1)
      bool first = true;     // nothing to observe before the very first plane
      for (int kl = 0; kl < my_units && ok; ++kl) {
        // virtual input planes 0 .. D0+1: the last two carry no MMAs, they only keep the barrier phases in step
        for (int xi = 0; xi < D0 + 2 && ok; ++xi) {
          const bool real = xi < D0;
          if (real) {
            long long tw0 = 0;
            if (tr) tw0 = clock64();
            ok = rs_wait(PFULL(slot), pphase, a.error_flag, 43, abort_s);
            if (tr) wait_p += clock64() - tw0;
            if (!ok) break;
            tc_fence_after();
          }
          const uint32_t wrot = (w16 + rot * WROT16) | b_lbo;
          uint32_t arow = ((ring16 + (uint32_t)slot * plane16) | a_lbo) + (uint32_t)iw * W;  // A descriptor of this issuer's next row
          const uint32_t astep = (uint32_t)ni * W;
          // Barrier operations cost 100-300 clk of this thread's time each (measured: a loop of nothing but two polls per
          // row ran at 400 clk / row), so both directions work on chunks of four rows: one "cleared" poll when the issuer
          // enters a chunk of output rows, one commit ("multiplied") when it leaves a chunk of input rows.
          // Input row r = i - 1 feeds output rows r-1 (k1 = 2), r (k1 = 1), r+1 (k1 = 0) = B rows [0,RP) [RP,2RP) [2RP,3RP);
          // the first / last two rows of the strip use the trailing / leading part of the same B tile.
          if (ni == 1) {
            // Single issuer (default).
#define SGM_RS_MMA3(AROW, B0, ID, COL)                                                       \
  do {                                                                                       \
    tc_mma(tmem_base + (COL), d_hi | (AROW), d_hi | (B0), (ID), 1u);                         \
    tc_mma(tmem_base + (COL), d_hi | ((AROW) + 1u), d_hi | ((B0) + 2u * NB), (ID), 1u);      \
    tc_mma(tmem_base + (COL), d_hi | ((AROW) + 2u), d_hi | ((B0) + 4u * NB), (ID), 1u);      \
  } while (0)
            // chunk by chunk (four input rows = one "cleared" poll + one commit); interior chunks -- four uniform middle
            // rows -- are straight-line code: 12 MMAs whose descriptors differ by constants
            const int last_chunk = (H1 - 1) >> 2;
            for (int c = 0; c <= last_chunk; ++c) {
              const int i0 = 4 * c;
              if (!first && i0 < t1) {  // entering the chunk of output rows i0 .. i0+3
                if (!rs_wait(CLR(c), vpar, a.error_flag, 44, abort_s)) { ok = false; break; }
                tc_fence_after();
              }
              if (real) {
                const uint32_t ar = arow + (uint32_t)i0 * W;
                if (i0 >= 2 && i0 + 3 <= t1 - 1) {
                  const uint32_t col = (uint32_t)(i0 - 2) * RP;
                  SGM_RS_MMA3(ar, wrot, id9, col);
                  SGM_RS_MMA3(ar + W, wrot, id9, col + RP);
                  SGM_RS_MMA3(ar + 2u * W, wrot, id9, col + 2u * RP);
                  SGM_RS_MMA3(ar + 3u * W, wrot, id9, col + 3u * RP);
                } else {
                  const int i1 = min(i0 + 4, H1);
                  for (int i = i0; i < i1; ++i) {
                    const int r = i - 1;
                    const uint32_t col = (uint32_t)max(r - 1, 0) * RP;
                    const uint32_t b0 = wrot + (r < 0 ? 2u * RP : (r == 0 ? (uint32_t)RP : 0u));
                    const uint32_t id = (r < 0 || r == t1) ? id3 : ((r == 0 || r == t1 - 1) ? id6 : id9);
                    SGM_RS_MMA3(arow + (uint32_t)i * W, b0, id, col);
                  }
                }
              }
              tc_commit(FULL(c));
            }
            if (!ok) break;
#undef SGM_RS_MMA3
          } else {
            int next_commit = 0;
            for (int i = iw; i < H1 && ok; i += ni, arow += astep) {
              const int c = i >> 2;
              if (!first && (i & 3) == iw && 4 * c < t1) {  // entering a chunk: output rows 4c .. 4c+3 (i-2, i-1: previous chunk)
                long long tw0 = 0;
                if (tr) tw0 = clock64();
                ok = rs_wait(CLR(c), vpar, a.error_flag, 44, abort_s);
                if (tr) wait_c += clock64() - tw0;
                tc_fence_after();
              }
              if (ok && real && !no_mma) {
                const int r = i - 1;
                const uint32_t col = (uint32_t)max(r - 1, 0) * RP;
                const uint32_t b0 = wrot + (r < 0 ? 2u * RP : (r == 0 ? (uint32_t)RP : 0u));
                const uint32_t id = (r < 0 || r == t1) ? id3 : ((r == 0 || r == t1 - 1) ? id6 : id9);
                tc_mma(tmem_base + col, d_hi | arow, d_hi | b0, id, 1u);
                if (!one_mma) {
                  tc_mma(tmem_base + col, d_hi | (arow + 1u), d_hi | (b0 + 2u * NB), id, 1u);
                  tc_mma(tmem_base + col, d_hi | (arow + 2u), d_hi | (b0 + 4u * NB), id, 1u);
                }
              }
              if (ok && ((i & 3) == 4 - ni + iw || i + ni >= H1)) {  // this issuer's last row of the chunk
                tc_commit(FULL(c));
                next_commit = c + 1;
              }
            }
            for (int c = next_commit; c <= (H1 - 1) >> 2 && ok; ++c) tc_commit(FULL(c));  // chunks without a row of this issuer
          }
          if (real && ok) {
            tc_commit(PEMPTY(slot));  // this issuer's MMAs that read the plane have completed when this arrives
            if (++slot == R) slot = 0, pphase ^= 1u;
          }
          first = false;
          vpar ^= 1u;
          if (++rot == 3) rot = 0;
        }
      }
      if (tr && iw == 0) a.trace[2] = clock64(), a.trace[3] = wait_p, a.trace[4] = wait_c;
    }
    __syncwarp();
  } else if (warp < kRsEpiWarps && (warp & 3) < nq) {
    // ============================ epilogue warps: group = warp / 4 owns output rows o = group, group + 4, ... ============================
    const int egroup = warp >> 2, quarter = warp & 3;
    const uint32_t tlane = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const int d2 = quarter * 32 + lane;
    const bool valid2 = d2 < a.D[2];
    const bool weighted = a.pl_weighted != 0;
    const bool res = a.res_mode == 2;
    const long long plane_vox = (long long)a.D[1] * a.D[2];
    const long long cs = a.pl_cstride;
    const float w2 = (weighted && valid2) ? __ldg(a.imap2 + d2) : 1.f;
    const float floor_w = weighted ? a.imap_floor : 1.f;
    float bias_r[CR];
#pragma unroll
    for (int c = 0; c < CR; ++c) bias_r[c] = bias_s[c];
    // One lane polls, the warp follows through the shuffle; a short sleep between polls leaves the issue slots of the
    // scheduler to the issuer warps.
    auto warp_wait = [&](uint32_t bar, uint32_t parity, int code) {
      bool r = true;
      if (lane == 0) r = rs_wait(bar, parity, a.error_flag, code, abort_s, 32u);
      return (bool)__shfl_sync(0xffffffffu, (int)r, 0);
    };
    bool ok = true;
    int rslot = 0;
    uint32_t rphase = 0, fpar = 0, slot3 = 0;  // parity of the virtual plane counter, and the counter mod 3
    long long wait_f = 0;
    const long long te0 = tr ? clock64() : 0;
    for (int kl = 0; kl < my_units && ok; ++kl) {
      const int unit = (int)blockIdx.x + kl * (int)gridDim.x;
      const int n = unit / a.nstrips, strip = unit - n * a.nstrips;
      const int r0 = strip * t1;
      float* out_n = a.pl_out + (long long)n * a.pl_nstride + (long long)r0 * a.D[2] + d2;
      const uint8_t* res_lane = ring + (size_t)(a.W + d2 + 1) * 16;  // residual of output row 0: input row 1, position d2 + 1
      for (int v = -1; v <= D0 && ok; ++v) {
        // virtual output plane v (counter V = kl * (D0 + 2) + v + 1): slot V % 3, complete once virtual input plane
        // v + 1 has been multiplied ("multiplied" phase index == V)
        const uint32_t pcol = slot3 * CP;
        const bool real = v >= 0 && v < D0;
        if (real && res) {
          // generic-proxy reads of TMA-written data: observe the plane's own barrier (it completed long ago)
          ok = warp_wait(PFULL(rslot), rphase, 45);
          if (!ok) break;
        }
        float g0 = 1.f;
        if (real && weighted) g0 = __ldg(a.imap0 + v);
        float* out_v = out_n + (long long)v * plane_vox;
        const uint8_t* res_v = res_lane + (size_t)rslot * plane_bytes;
        // A step is a chain of dependent long-latency operations (barrier poll, TMEM load, TMEM store, shared and
        // global loads: ~1000 clk measured, whatever the amount of arithmetic), so each step handles TWO adjacent output
        // rows: group g owns rows 2g, 2g+1, then 2g+8, 2g+9.
        for (int oa = 2 * egroup; oa < t1; oa += 8) {
          const bool has_b = oa + 1 < t1;
          // input rows o, o+1, o+2 feed output row o: complete when the chunk of input row o+2 has been committed
          long long tw0 = 0;
          if (tr) tw0 = clock64();
          ok = warp_wait(FULL((oa + (has_b ? 3 : 2)) >> 2), fpar, 46);
          if (tr) wait_f += clock64() - tw0;
          if (!ok) break;
          tc_fence_after();
          const uint32_t taddr = tlane + (uint32_t)(oa * RP) + pcol;
          uint32_t acc[2][CP];
          rs_ld<CP>(taddr, acc[0]);
          if (has_b) rs_ld<CP>(taddr + RP, acc[1]);
          tc_ld_wait();
          rs_st_zero<CP>(taddr);  // the plane three steps later accumulates into these columns again
          if (has_b) rs_st_zero<CP>(taddr + RP);
          rs_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(CLR(oa >> 2));
            if (has_b) mbar_arrive(CLR(oa >> 2));  // oa is even: both rows are in the same chunk
          }
          if (real && valid2) {
            float g1[2] = {1.f, 1.f};
            if (weighted) {
              g1[0] = __ldg(a.imap1 + min(r0 + oa, a.D[1] - 1));
              g1[1] = __ldg(a.imap1 + min(r0 + oa + 1, a.D[1] - 1));
            }
            uint4 rres[2][(CR + 7) / 8];
            if (res) {
#pragma unroll
              for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int pc = 0; pc < (CR + 7) / 8; ++pc)
                  rres[h][pc] = *reinterpret_cast<const uint4*>(res_v + (size_t)((oa + h) * a.W) * 16 + (size_t)pc * a.GS * 16);
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              if (h == 1 && !has_b) break;
              if (r0 + oa + h >= a.D[1]) break;
              float val[CP];
#pragma unroll
              for (int c = 0; c < CR; ++c) val[c] = __uint_as_float(acc[h][c]) + bias_r[c];
              if (res) {
#pragma unroll
                for (int pc = 0; pc < (CR + 7) / 8; ++pc) {
                  float rr[8];
                  unpack8(rres[h][pc], rr);
#pragma unroll
                  for (int c = 0; c < 8; ++c)
                    if (8 * pc + c < CR) val[8 * pc + c] += rr[c];
                }
              }
              // importance weight with MONAI's roundings; the unweighted head multiplies by exactly 1.0
              float imw = 1.f;
              if (weighted) imw = fmaxf(__fmul_rn(__fmul_rn(g0, g1[h]), w2), floor_w);
              float* dst = out_v + (long long)(oa + h) * a.D[2];
#pragma unroll
              for (int c = 0; c < CR; ++c, dst += cs) __stcs(dst, __fmul_rn(val[c], imw));
            }
          }
        }
        if (real && res && ok) {
          __syncwarp();
          if (lane == 0) mbar_arrive(PEMPTY(rslot));  // this warp no longer reads the plane
        }
        if (real && ++rslot == R) rslot = 0, rphase ^= 1u;  // ring entry / parity of the next real plane
        fpar ^= 1u;
        if (++slot3 == 3) slot3 = 0;
      }
    }
    if (tr && warp == 0 && lane == 0) a.trace[6] = wait_f, a.trace[7] = clock64() - te0;
  }

  tc_fence_before();
  __syncthreads();
  if (warp == PW) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
  if (tr && tid == 0) a.trace[5] = clock64();
}

inline uint16_t f2bf(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7f800000u) != 0x7f800000u) u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}

inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

struct RsPlan {
  int key[4];
  RsArgs args;
  int smem_bytes, grid;
};

int rs_sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    if (const char* env = getenv("SGM_RESERVE_SMS")) n = std::max(8, n - atoi(env));
  }
  return n;
}

int rs_fixed_smem(int cp) {
  return round_up(rs_w_bytes(cp), 128) + kRsTailPad + 32 * 4 + (1 + 2 * kRsRingMax + (kRsRowsMax + 2) + kRsRowsMax) * 8 + 16 + 128;  // + TMEM slot, abort flag, slack
}

int rs_plan(int cp, const TcIO& io, RsPlan& pl) {
  RsArgs& a = pl.args;
  memset(&a, 0, sizeof(a));
  for (int i = 0; i < 3; ++i) a.D[i] = io.od[i];
  a.W = a.D[2] + 2;
  a.nq = ceil_div(a.D[2], 32);
  const int nsm = rs_sm_count();
  const int fixed = rs_fixed_smem(cp);
  const double t_row = 3.0 * (34.0 + 0.36 * 3 * rs_rp(cp));
  // TMEM: the widest accumulate of the strip must end inside the 512 columns
  int t1_max = kRsRowsMax;
  while (t1_max > 3 && t1_max * rs_rp(cp) > 512) --t1_max;
  int t1_min = 3;
  if (const char* env = getenv("SGM_RS_T1")) t1_min = t1_max = std::max(3, std::min(t1_max, atoi(env)));  // diagnostic: force the strip height
  double best = 1e30;
  int bt1 = 0, bR = 0;
  for (int t1 = t1_min; t1 <= std::min(t1_max, std::max(3, a.D[1])); ++t1) {
    const int gs = round_up((t1 + 2) * a.W, 8);
    const int plane_bytes = 2 * gs * 16;
    const int R = std::min(kRsRingMax, (kRsSmemMax - fixed) / plane_bytes);
    if (R < 3) continue;
    const long long units = (long long)ceil_div(a.D[1], t1) * io.n;
    const double waves = (double)((units + nsm - 1) / nsm);
    // a ring of three planes (residual / multiplied / loading) leaves the TMA no slack: slight preference for four
    const double cost = waves * ((double)a.D[0] * (t1 + 2) * t_row + 6000.0) * (R < 4 ? 1.05 : 1.0);
    if (cost < best) best = cost, bt1 = t1, bR = R;
  }
  SGM_REQUIRE(bt1 > 0, SGM_ERR_UNSUPPORTED, "rs_plan: no strip height fits shared memory");
  a.t1 = bt1, a.H1 = bt1 + 2, a.R = std::min(bR, 4);
  a.GS = round_up(a.H1 * a.W, 8);
  a.issuers = 1;
  if (const char* env = getenv("SGM_RS_ISSUERS")) a.issuers = atoi(env) == 2 ? 2 : 1;
  a.nstrips = ceil_div(a.D[1], bt1);
  a.nunits = a.nstrips * io.n;
  pl.smem_bytes = fixed + a.R * 2 * a.GS * 16;
  pl.grid = std::min(nsm, a.nunits);
  return SGM_OK;
}

}  // namespace

// Packs [rot][k2][kchunk][NB][8]: row (2 - k1) * 3CP + p * CP + co holds W[co][ci][k0 = (rot + 2 - p) mod 3][k1][k2].
int rs_pack(const sgm_conv_desc& d, TcConv* c) {
  c->rs_cp = 0;
  if (getenv("SGM_NO_RS")) return SGM_OK;
  if (d.kind != SGM_KIND_CONV || d.kernel != 3 || d.stride != 1 || c->flat0 || c->mode != MODE_S1) return SGM_OK;
  if (d.cout > 10 || d.cin > 16 || c->cgin != 2) return SGM_OK;
  const int cp = 10, NB = rs_nb(cp);
  std::vector<uint16_t> w((size_t)rs_w_bytes(cp) / 2, 0);
  for (int rot = 0; rot < 3; ++rot)
    for (int k2 = 0; k2 < 3; ++k2)
      for (int kc = 0; kc < 2; ++kc)
        for (int k1 = 0; k1 < 3; ++k1)
          for (int p = 0; p < 3; ++p) {
            const int k0 = (rot + 2 - p) % 3;
            for (int co = 0; co < d.cout; ++co)
              for (int k8 = 0; k8 < 8; ++k8) {
                const int ci = kc * 8 + k8;
                if (ci >= d.cin) continue;
                const int tap = (k0 * 3 + k1) * 3 + k2;
                const int row = (2 - k1) * rs_rp(cp) + p * cp + co;
                w[((((size_t)rot * 3 + k2) * 2 + kc) * NB + row) * 8 + k8] = f2bf(d.weight[((size_t)co * d.cin + ci) * 27 + tap]);
              }
          }
  if (cudaMalloc(&c->rs_w, w.size() * 2) != cudaSuccess) {
    set_error("rs_pack: cudaMalloc failed: %s", cudaGetErrorString(cudaGetLastError()));
    return SGM_ERR_CUDA;
  }
  SGM_CUDA_CHECK(cudaMemcpy(c->rs_w, w.data(), w.size() * 2, cudaMemcpyHostToDevice));
  c->rs_cp = cp;
  c->rs_plan_cache = new std::vector<RsPlan>();
  return SGM_OK;
}

void rs_free(TcConv* c) {
  if (c->rs_w) cudaFree(c->rs_w);
  c->rs_w = nullptr;
  delete reinterpret_cast<std::vector<RsPlan>*>(c->rs_plan_cache);
  c->rs_plan_cache = nullptr;
}

bool rs_applicable(const TcConv& c, const TcIO& io) {
  if (!c.rs_cp || !c.rs_w || !tma_available()) return false;
  if (io.in1 || io.cg1 || io.outB || io.cg0 != 2) return false;
  if (io.out_kind != OUT_PLANAR || c.actA) return false;  // planar, conv-only (MONAI's head); the read-modify-write blend stays on conv_ps
  if (io.res && io.res != io.in0) return false;            // identity residual only
  for (int i = 0; i < 3; ++i)
    if (io.id[i] != io.od[i]) return false;
  if (io.od[2] > 126 || io.od[2] < 33 || io.od[1] < 3) return false;  // one input row (+ halo) = one 128-lane tile, mostly filled
  if ((long long)io.od[0] * io.od[1] * io.od[2] >= (1LL << 31)) return false;
  return true;
}

int rs_launch(const TcConv& c, const TcIO& io, int* error_flag_dev, cudaStream_t st) {
  auto* plans = reinterpret_cast<std::vector<RsPlan>*>(c.rs_plan_cache);
  const int key[4] = {io.od[0], io.od[1], io.od[2], io.n};
  const RsPlan* pe = nullptr;
  for (auto& e : *plans)
    if (memcmp(e.key, key, sizeof(key)) == 0) pe = &e;
  if (!pe) {
    RsPlan e;
    memcpy(e.key, key, sizeof(key));
    int rc = rs_plan(c.rs_cp, io, e);
    if (rc) return rc;
    plans->push_back(e);
    pe = &plans->back();
  }
  RsArgs a = pe->args;
  a.c_real = c.c_real;
  a.w = c.rs_w, a.bias = c.bias;
  a.res_mode = io.res ? 2 : 0;
  a.pl_weighted = io.pl_weighted;
  a.pl_out = io.pl_out, a.pl_cstride = io.pl_cstride, a.pl_nstride = io.pl_nstride;
  a.imap0 = io.imap[0], a.imap1 = io.imap[1], a.imap2 = io.imap[2], a.imap_floor = io.imap_floor;
  a.error_flag = error_flag_dev;
  static const bool dbg = getenv("SGM_DEBUG") != nullptr;
  if (dbg)
    fprintf(stderr, "[rs_launch] CP=%d D=(%d,%d,%d) n=%d t1=%d strips=%d units=%d grid=%d smem=%d R=%d GS=%d res=%d weighted=%d\n",
            c.rs_cp, a.D[0], a.D[1], a.D[2], io.n, a.t1, a.nstrips, a.nunits, pe->grid, pe->smem_bytes, a.R, a.GS,
            a.res_mode, a.pl_weighted);
  if (const char* env = getenv("SGM_RS_ABL")) a.abl = atoi(env);
  static const bool trace_on = getenv("SGM_TRACE") != nullptr;
  static long long* trace_dev = nullptr;
  if (trace_on) {
    if (!trace_dev) cudaMalloc(&trace_dev, 32 * sizeof(long long));
    cudaMemsetAsync(trace_dev, 0, 32 * sizeof(long long), st);
    a.trace = trace_dev;
  }
  CUtensorMap tm;
  const int box[3] = {1, a.H1, a.W};
  const int par[3] = {1, 1, 1};  // rank-5 map {8 ch, D2, D1, D0, n * groups}: a row of W positions exceeds the 256-element box of the merged form
  int rc = make_brick_map(&tm, io.in0, io.n * io.cg0, io.id, box, par);
  if (rc) return rc;
  rc = SGM_ERR_UNSUPPORTED;
#define SGM_RS_CASE(CRV)                                                                                              \
  if (a.c_real == CRV) {                                                                                              \
    static DeviceOnce attr_once;                                                                                      \
    if (attr_once.first()) {                                                                                          \
      SGM_CUDA_CHECK(cudaFuncSetAttribute(rs_conv_kernel<10, CRV>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRsSmemMax)); \
    }                                                                                                                 \
    rs_conv_kernel<10, CRV><<<pe->grid, kRsThreads, pe->smem_bytes, st>>>(a, tm);                                     \
    rc = SGM_OK;                                                                                                      \
  }
  SGM_RS_CASE(1) SGM_RS_CASE(2) SGM_RS_CASE(3) SGM_RS_CASE(4) SGM_RS_CASE(5) SGM_RS_CASE(6) SGM_RS_CASE(7) SGM_RS_CASE(8)
  SGM_RS_CASE(9) SGM_RS_CASE(10)
#undef SGM_RS_CASE
  SGM_REQUIRE(rc == SGM_OK, SGM_ERR_UNSUPPORTED, "rs_launch: no instantiation for %d output channels", a.c_real);
  SGM_CUDA_CHECK(cudaGetLastError());
  if (trace_on) {
    long long t[32];
    cudaStreamSynchronize(st);
    cudaMemcpy(t, trace_dev, sizeof(t), cudaMemcpyDeviceToHost);
    const double units = (double)((a.nunits - 1) / pe->grid + 1), rows = units * a.D[0] * a.H1;
    fprintf(stderr,
            "[rs trace] D=(%d,%d,%d) n=%d t1=%d RC=%d units/cta=%.0f | issuer: start %lld end %lld (%.0f clk / input row), waited %lld "
            "for planes, %lld for cleared rows | epilogue warp 0: %lld clk, waited %lld for MMAs | kernel %lld clk\n",
            a.D[0], a.D[1], a.D[2], io.n, a.t1, 4, units, t[1] - t[0], t[2] - t[0], (double)(t[2] - t[1]) / rows, t[3], t[4],
            t[7], t[6], t[5] - t[0]);
  }
  return SGM_OK;
}

}  // namespace tc
}  // namespace sgm
