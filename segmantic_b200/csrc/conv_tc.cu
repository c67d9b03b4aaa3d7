// tcgen05 implicit-GEMM convolutions for the MONAI UNet residual units (bf16 in, fp32 TMEM accumulate).
//
// Data layout.  Activations are CG8 bf16: [n][cg][d0][d1][d2][8 ch] -- one voxel's 8-channel group is
// 16 bytes, and 8 consecutive d2 voxels of one group are 128 contiguous bytes: exactly one K-major
// SWIZZLE_NONE "core matrix" (8 rows x 16 B) of a tcgen05 shared-memory descriptor.
//
// Algorithm.  A CTA stages the halo brick of its output tile ONCE in shared memory, as
// [slab][cg][h0][h1][h2] x 16 B.  GEMM row m is a *padded linear* brick position p, so the A operand of
// filter tap (k0,k1,k2) for 128 consecutive rows is the same brick read at p + shift(k): a descriptor
// with start address + shift*16, SBO = 128 B (next 8 rows), LBO = one channel-group slab.  No im2col,
// no per-tap reload: every tap / channel-pair is one MMA (M=128, N=Cout, K=16) straight from the brick.
// Rows that fall in the halo are computed and discarded (tile shapes are chosen to keep them ~25 %).
//   stride 2 (down path): the brick is loaded as 8 parity slabs (space-to-depth), taps pick slab+shift;
//                         the residual-branch conv of the block reads the same A and is fused as extra N.
//   transposed stride 2:  rows are INPUT voxels; the 8 output parity classes are 8 accumulators with
//                         1/2/2/2/4/4/4/8 taps; A is the channel concat of skip and sub-network
//                         tensors (two base pointers -- torch.cat is never materialised).
// Warp roles (192 threads): warps 0-3 epilogue (TMEM -> regs -> bias/PReLU/residual -> bf16 CG8 store,
// or fp32 importance-weighted accumulate into the volume for the head), warp 4 MMA issuer (one
// thread), warp 5 weight producer (cp.async.bulk ring, mbarrier complete_tx).  The brick itself is
// gathered by all threads with zero-filling cp.async (window-border zero padding for free).
// Accumulators are double-buffered in TMEM so the epilogue of chunk i overlaps the MMAs of chunk i+1;
// two CTAs per SM overlap brick loads with the other CTA's MMAs.
#include "conv_tc.cuh"
#include "tc_ptx.cuh"

#include <cuda.h>  // CUtensorMap (types only; the encoder is resolved through cudaGetDriverEntryPoint)

#include <algorithm>
#include <stdlib.h>
#include <string.h>

namespace sgm {
namespace tc {

namespace {

constexpr int kThreads = 192;
constexpr int kSmemBudget = 110 * 1024;   // two CTAs per SM

// CTAs per SM the launch plan aims for (SGM_TC_CTAS = 2 or 3): the one-CTA-per-brick kernel is latency bound (TMEM
// allocation, brick TMA, weight ring, 27-108 MMAs, epilogue in sequence), so what counts is how many bricks overlap.
static int target_ctas() {
  static int v = 0;
  if (!v) {
    v = 3;
    if (const char* env = getenv("SGM_TC_CTAS")) v = atoi(env) == 2 ? 2 : 3;
  }
  return v;
}
using namespace tcptx;

struct KArgs {
  const __nv_bfloat16* in0;
  const __nv_bfloat16* in1;
  int cg0, cg1, cgin;
  int id[3], od[3], rd[3];
  int mode;
  int t[3], H[3], lo[3], nt[3];
  int ibase_mul[3];  // input coord = org*omul + ioff + imul*h + r  (see loader)
  int imul[3], ioff[3], par[3];
  int P, nslab;
  int row_first, ntiles, tpc, nchunks, nbuf, ncls, cols_per_buf, tmem_cols;
  int N, nblk, G, ngroups, nstages, resident;
  int achunk;         // > 0: the brick is streamed in chunks of `achunk` channel groups through a ring (deep layers)
  int astages, astage_units, nkc;
  int mmaN;           // MMA N: N, or ncls * N when the output parity classes of a transposed conv are folded into one MMA
  int blk_off;        // offset of this conv's K-block descriptors in c_blk
  const __nv_bfloat16* w;
  const float* bias;
  __nv_bfloat16* outA;
  int cgA;
  __nv_bfloat16* outB;
  int cgB;
  int segA_cg, actA;
  float alphaA;
  const __nv_bfloat16* res;
  int res_mode;       // 0 none, 1 global CG8 tensor, 2 identity: centre of the shared-memory brick
  int pl_weighted;    // OUT_PLANAR: multiply by the window importance map (deferred blend)
  int out_kind;
  float* pl_out;
  long long pl_cstride, pl_nstride;
  int ad0, ad1, ad2;
  int wo[3];
  const float* imap0;
  const float* imap1;
  const float* imap2;
  float imap_floor;
  int c_real;
  int use_tma;        // brick loaded by TMA (cp.async.bulk.tensor) instead of the cp.async gather
  int box_bytes;      // bytes one TMA box (one channel group of the brick) deposits
  int a_units;        // 16-byte units reserved for the A region
  int w_stage_bytes;
  int* error_flag;
  long long* trace;   // optional (SGM_TRACE): clock64 stamps of CTA 0's phases
};


// ------------------------------------------------------------------------------------------ kernel
// K-block descriptors of every distinct (mode, kernel, channel-pair count) live in constant memory so
// the MMA warp reads them through the uniform datapath (no per-lane waterfall around tcgen05.mma).
constexpr int kBlkConst = 12288;
__constant__ uint32_t c_blk[kBlkConst];


// ACH: the brick is streamed in channel chunks (deep layers, one CTA per SM); compiled out of the common path
// MINB = CTAs per SM the register allocation aims for: 2 (168 registers) or 3 (96 registers, a few spills in the
// epilogue) -- chosen per layer by the launch plan (tc_launch)
template <bool ACH, int MINB = 2>
__global__ void __launch_bounds__(kThreads, ACH ? 1 : MINB)
tc_conv_kernel(const KArgs a, const __grid_constant__ CUtensorMap tmap0, const __grid_constant__ CUtensorMap tmap1) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int coblk = blockIdx.y, n = blockIdx.z;
  int tile_lin = blockIdx.x;
  const int b2 = tile_lin % a.nt[2];
  tile_lin /= a.nt[2];
  const int b1 = tile_lin % a.nt[1];
  const int b0 = tile_lin / a.nt[1];
  const int org[3] = {b0 * a.t[0], b1 * a.t[1], b2 * a.t[2]};  // tile origin in row space
  const bool tr = a.trace != nullptr && blockIdx.x == gridDim.x / 2 && blockIdx.y == 0 && blockIdx.z == 0;
#define SGM_TRACE(slot) do { if (tr && lane == 0) a.trace[slot] = clock64(); } while (0)
  if (warp == 0) SGM_TRACE(0);

  uint8_t* a_smem = smem;
  const int a_bytes = a.a_units * 16;
  uint8_t* w_smem = smem + a_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(w_smem + a.nstages * a.w_stage_bytes);
  // bars: [0,nstages) wfull, [nstages,2nstages) wempty, then tfull[2], tempty[2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * a.nstages + 13);
  uint2* blk_tab = reinterpret_cast<uint2*>(bars + 2 * a.nstages + 14);  // per K block: {A offset, column | chunk << 16 | overwrite << 31}
  const uint32_t bar0 = smem_u32(bars);
  auto WFULL = [&](int s) { return bar0 + 8u * s; };
  auto WEMPTY = [&](int s) { return bar0 + 8u * (a.nstages + s); };
  auto TFULL = [&](int b) { return bar0 + 8u * (2 * a.nstages + b); };
  auto TEMPTY = [&](int b) { return bar0 + 8u * (2 * a.nstages + 2 + b); };
  const uint32_t ABAR = bar0 + 8u * (2 * a.nstages + 4);  // brick landed (TMA path)
  auto AFULL = [&](int s) { return bar0 + 8u * (2 * a.nstages + 5 + s); };   // channel-chunk ring (a.achunk > 0)
  auto AEMPTY = [&](int s) { return bar0 + 8u * (2 * a.nstages + 9 + s); };

  if (warp == 4 && lane == 0) {
    for (int s = 0; s < a.nstages; ++s) {
      mbar_init(WFULL(s), 1);
      mbar_init(WEMPTY(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(TFULL(b), 1);
      mbar_init(TEMPTY(b), 128);
    }
    mbar_init(ABAR, 1);
    for (int st = 0; st < 4; ++st) {
      mbar_init(AFULL(st), 1);
      mbar_init(AEMPTY(st), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)a.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }

  // ---- gather the halo brick: [slab][cg][h0][h1][h2] x 16 B, zero fill outside the window.
  // One warp per brick row (fixed slab, cg, h0, h1), lanes along h2 -> no per-element div/mod and
  // 16-byte requests that are contiguous in global memory for stride-1 bricks.
  if (!a.use_tma) {
    const long long ivox = (long long)a.id[0] * a.id[1] * a.id[2];
    const uint32_t a_base = smem_u32(a_smem);
    const int ib0 = org[0] * a.ibase_mul[0] + a.ioff[0], ib1 = org[1] * a.ibase_mul[1] + a.ioff[1],
              ib2 = org[2] * a.ibase_mul[2] + a.ioff[2];
    const int nrows = a.nslab * a.cgin * a.H[0] * a.H[1];
    // row = ((slab*cgin + cg)*H0 + h0)*H1 + h1, advanced incrementally (no per-row division)
    int row = warp;
    int h1 = row % a.H[1];
    int pl = row / a.H[1];
    int h0 = pl % a.H[0];
    int sc = pl / a.H[0];
    int cg = sc % a.cgin, slab = sc / a.cgin;
    while (row < nrows) {
      int bits = slab;  // parity bits, least significant = axis 2 (only axes with par==2 consume a bit)
      const int r2 = a.par[2] == 2 ? (bits & 1) : 0;
      bits >>= (a.par[2] == 2);
      const int r1 = a.par[1] == 2 ? (bits & 1) : 0;
      bits >>= (a.par[1] == 2);
      const int r0 = a.par[0] == 2 ? (bits & 1) : 0;
      const int i0 = ib0 + a.imul[0] * h0 + r0, i1 = ib1 + a.imul[1] * h1 + r1;
      const bool ok01 = i0 >= 0 && i0 < a.id[0] && i1 >= 0 && i1 < a.id[1];
      const __nv_bfloat16* base = (cg < a.cg0) ? a.in0 + ((long long)n * a.cg0 + cg) * ivox * 8
                                               : a.in1 + ((long long)n * a.cg1 + (cg - a.cg0)) * ivox * 8;
      const long long rowoff = ((long long)i0 * a.id[1] + i1) * a.id[2];
      const uint32_t dst_row = a_base + (uint32_t)((slab * a.cgin + cg) * a.P + (h0 * a.H[1] + h1) * a.H[2]) * 16u;
      for (int h2 = lane; h2 < a.H[2]; h2 += 32) {
        const int i2 = ib2 + a.imul[2] * h2 + r2;
        const bool ok = ok01 && i2 >= 0 && i2 < a.id[2];
        const __nv_bfloat16* src = ok ? base + (rowoff + i2) * 8 : a.in0;
        cp_async16(dst_row + (uint32_t)h2 * 16u, src, ok ? 16u : 0u);
      }
      row += kThreads / 32;
      h1 += kThreads / 32;
      while (h1 >= a.H[1]) {
        h1 -= a.H[1];
        if (++h0 == a.H[0]) {
          h0 = 0;
          if (++cg == a.cgin) cg = 0, ++slab;
        }
      }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
  }
  // K-block table: everything the issuer needs per MMA that does not depend on the tile, decoded once
  // (decoding the descriptor words inside the issue loop cost ~200 clk per MMA: the loop was issue-bound)
  for (int b = tid; b < a.nblk; b += kThreads) {
    const uint32_t d = c_blk[a.blk_off + b];
    const int s0 = (int)(d & 3u) - 1, s1 = (int)((d >> 2) & 3u) - 1, s2 = (int)((d >> 4) & 3u) - 1;
    const int slab = (d >> 6) & 7u, cls = (d >> 9) & 7u, cgpair = (d >> 16) & 0xffu;
    int a16 = a.row_first + s0 * a.H[1] * a.H[2] + s1 * a.H[2] + s2;
    uint32_t kc = 0;
    if (ACH) {  // offset inside the ring stage of the block's channel chunk
      const int cpc = a.achunk >> 1;
      kc = (uint32_t)(cgpair / cpc);
      a16 += (cgpair % cpc) * 2 * a.P;
    } else {
      a16 += (slab * a.cgin + cgpair * 2) * a.P;
    }
    blk_tab[b] = make_uint2((uint32_t)a16, (uint32_t)(cls * a.N) | (kc << 16) | (((d >> 12) & 1u) << 31));
  }
  // zero the tail the last M tile may read (keeps garbage rows finite; they are discarded anyway)
  for (int it = (ACH ? 0 : a.nslab * a.cgin * a.P) + tid; it < a.a_units; it += kThreads)
    *reinterpret_cast<uint4*>(a_smem + (size_t)it * 16) = make_uint4(0, 0, 0, 0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> tensor-core reads
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 0) SGM_TRACE(1);

  const int N = a.N;
  if (warp == 5) {
    // ================= weight producer: cp.async.bulk ring =================
    if (lane == 0) {
      if (a.use_tma && !ACH) {
        // halo brick by TMA: one 4-D box {H2*8 elements, H1, H0, 1 channel group} per group; coordinates may
        // be negative / past the extent -> hardware zero fill == the conv's zero padding at the window border
        mbar_expect_tx(ABAR, (uint32_t)(a.nslab * a.cgin * a.box_bytes));
        const uint32_t a_base = smem_u32(a_smem);
        const int c0 = org[2] * a.ibase_mul[2] + a.ioff[2], c1 = org[1] * a.ibase_mul[1] + a.ioff[1],
                  c2 = org[0] * a.ibase_mul[0] + a.ioff[0];
        if (a.mode == MODE_S2) {
          // stride 2: one rank-5 box per (parity slab, channel group), element strides 2 along the split axes
          for (int slab = 0; slab < a.nslab; ++slab) {
            int bits = slab;
            const int r2 = a.par[2] == 2 ? (bits & 1) : 0;
            bits >>= (a.par[2] == 2);
            const int r1 = a.par[1] == 2 ? (bits & 1) : 0;
            bits >>= (a.par[1] == 2);
            const int r0 = a.par[0] == 2 ? (bits & 1) : 0;
            for (int cg = 0; cg < a.cgin; ++cg)
              tma_load_5d(a_base + (uint32_t)((slab * a.cgin + cg) * a.P) * 16u, &tmap0, 0, c0 + r2, c1 + r1, c2 + r0,
                          n * a.cg0 + cg, ABAR);
          }
        } else {
          for (int cg = 0; cg < a.cgin; ++cg) {
            const bool first_src = cg < a.cg0;
            tma_load_4d(a_base + (uint32_t)(cg * a.P) * 16u, first_src ? &tmap0 : &tmap1, c0 * 8, c1, c2,
                        first_src ? n * a.cg0 + cg : n * a.cg1 + (cg - a.cg0), ABAR);
          }
        }
      }
      const int NB = a.mmaN;
      const __nv_bfloat16* wsrc = a.w + (size_t)coblk * a.nblk * NB * 16;
      const int total = a.resident ? a.ngroups : a.nchunks * a.ngroups;
      for (int it = 0; it < total; ++it) {
        const int g = it % a.ngroups, stage = it % a.nstages;
        if (it >= a.nstages) {
          const uint32_t ph = (uint32_t)(it / a.nstages) & 1u;
          if (!mbar_wait(WEMPTY(stage), ph ^ 1u, a.error_flag, 1)) break;
        }
        const int nb = min(a.G, a.nblk - g * a.G);
        const uint32_t bytes = (uint32_t)nb * NB * 32u;
        mbar_expect_tx(WFULL(stage), bytes);
        bulk_g2s(smem_u32(w_smem + (size_t)stage * a.w_stage_bytes), wsrc + (size_t)g * a.G * NB * 16, bytes,
                 WFULL(stage));
      }
    }
  } else if (warp == 4) {
    // ================= MMA issuer: ONE elected lane runs the whole issue loop ======================
    // (elect.sync tells the compiler exactly one lane is live, so descriptors move to uniform
    // registers with plain R2UR -- no per-MMA election / reconvergence and no waterfall loops.)
    if (elect_one()) {
      // instruction descriptor: D=f32, A=B=bf16, K-major both, N>>3 at [17,23), M>>4 at [24,29)
      const int NB = a.mmaN;
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NB >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t a_base16 = smem_u32(a_smem) >> 4;
      const uint32_t w_base16 = smem_u32(w_smem) >> 4;
      const uint32_t desc_hi = 8u | (1u << 14);                 // SBO = 128 B, version 1
      const uint32_t a_lbo = ((uint32_t)a.P & 0x3FFFu) << 16;   // next channel group of the brick
      const uint32_t b_lbo = ((uint32_t)NB & 0x3FFFu) << 16;    // next 8 input channels of the filter block
      const uint32_t tile_cols = (uint32_t)(a.ncls * N);
      if (a.use_tma && !ACH) mbar_wait_or_trap(ABAR, 0u, a.error_flag, 5);
      int cur_kc = -1;  // channel chunk whose ring stage the issuer currently reads
      if (tr) a.trace[2] = clock64();
      for (int chunk = 0; chunk < a.nchunks; ++chunk) {
        const int buf = a.nbuf == 2 ? (chunk & 1) : 0;
        const int use = a.nbuf == 2 ? (chunk >> 1) : chunk;
        if (use > 0) mbar_wait_or_trap(TEMPTY(buf), (uint32_t)(use - 1) & 1u, a.error_flag, 2);
        tc_fence_after();
        const int tiles_here = min(a.tpc, a.ntiles - chunk * a.tpc);
        for (int g = 0; g < a.ngroups; ++g) {
          int stage = g;
          if (!a.resident || chunk == 0) {
            const int it = a.resident ? g : chunk * a.ngroups + g;
            stage = it % a.nstages;
            mbar_wait_or_trap(WFULL(stage), (uint32_t)(it / a.nstages) & 1u, a.error_flag, 3);
            tc_fence_after();
            if (tr && chunk == 0 && g == 0) a.trace[3] = clock64();
          }
          const int bend = min(a.nblk, (g + 1) * a.G);
          uint32_t b_lo = (w_base16 + (uint32_t)(stage * a.w_stage_bytes) / 16u) | b_lbo;
          uint32_t a_chunk = (a_base16 + (uint32_t)(chunk * a.tpc * 128)) | a_lbo;  // < 2^14 units: no masking
          const uint32_t col_chunk = tmem_base + (uint32_t)(buf * a.cols_per_buf);
#pragma unroll 2
          for (int b = g * a.G; b < bend; ++b, b_lo += (uint32_t)(NB * 2)) {
            const uint2 e = blk_tab[b];
            if (ACH) {
              const int kc = (int)((e.y >> 16) & 0xffu);
              if (kc != cur_kc) {  // next channel chunk: release the stage just read, wait for the new one
                if (cur_kc >= 0) tc_commit(AEMPTY(cur_kc % a.astages));
                mbar_wait_or_trap(AFULL(kc % a.astages), (uint32_t)(kc / a.astages) & 1u, a.error_flag, 6);
                tc_fence_after();
                cur_kc = kc;
              }
              a_chunk = (a_base16 + (uint32_t)((kc % a.astages) * a.astage_units)) | a_lbo;
            }
            const uint32_t acc = (e.y >> 31) ? 0u : 1u;
            const uint64_t bdesc = ((uint64_t)desc_hi << 32) | b_lo;
            uint32_t a_lo = a_chunk + e.x;
            uint32_t col = col_chunk + (e.y & 0xffffu);
#pragma unroll 4
            for (int t = 0; t < tiles_here; ++t, a_lo += 128u, col += tile_cols)
              tc_mma(col, ((uint64_t)desc_hi << 32) | a_lo, bdesc, idesc, acc);
          }
          if (!a.resident) tc_commit(WEMPTY(stage));
        }
        tc_commit(TFULL(buf));
        if (tr && chunk < 4) a.trace[4 + chunk] = clock64();  // all MMAs of the chunk issued
      }
    }
    __syncwarp();
  } else {
    // ================= epilogue warps 0..3: TMEM lanes 32*warp .. 32*warp+31 =================
    // Work items = (tile, class, 16-column piece) of the chunk, walked linearly so that the global
    // residual of item i+1 is in flight while item i is processed.  Identity residuals (up-path
    // residual units and the head: residual == the conv input) come from the shared-memory brick.
    if (ACH) {
      // deep layers: the brick does not fit shared memory with all its channels, so it is streamed in chunks of
      // `achunk` channel groups through a TMA ring while the accumulators of ALL tiles stay in TMEM.  The epilogue
      // warps idle during the main loop: lane 0 of warp 0 feeds the ring.
      if (warp == 0 && lane == 0) {
        const uint32_t a_base = smem_u32(a_smem);
        const int c0 = org[2] - a.lo[2], c1 = org[1] - a.lo[1], c2 = org[0] - a.lo[0];
        for (int kc = 0; kc < a.nkc; ++kc) {
          const int stage = kc % a.astages;
          if (kc >= a.astages) mbar_wait_or_trap(AEMPTY(stage), (uint32_t)(kc / a.astages - 1) & 1u, a.error_flag, 7);
          mbar_expect_tx(AFULL(stage), (uint32_t)(a.achunk * a.box_bytes));
          for (int i = 0; i < a.achunk; ++i) {
            const int cg = kc * a.achunk + i;
            const bool first_src = cg < a.cg0;
            tma_load_4d(a_base + (uint32_t)(stage * a.astage_units + i * a.P) * 16u, first_src ? &tmap0 : &tmap1, c0 * 8, c1,
                        c2, first_src ? n * a.cg0 + cg : n * a.cg1 + (cg - a.cg0), AFULL(stage));
          }
        }
      }
      __syncwarp();
    }
    const long long ovox = (long long)a.od[0] * a.od[1] * a.od[2];
    const int npiece = N / 16;
    bool ok = true;
    // Fast path (N == 16, one class, at most 4 tiles per chunk: the 16-channel layers and the head):
    // all per-tile geometry is computed BEFORE waiting for the tensor core, bias lives in registers.
    const bool fast = a.ncls == 1 && npiece == 1 && a.out_kind != OUT_BLEND && a.tpc <= 4;
    float bias_r[16];
    if (fast) {
#pragma unroll
      for (int c = 0; c < 16; ++c) bias_r[c] = __ldg(a.bias + coblk * N + c);
    }
    for (int chunk = 0; chunk < a.nchunks && ok; ++chunk) {
      const int buf = a.nbuf == 2 ? (chunk & 1) : 0;
      const int use = a.nbuf == 2 ? (chunk >> 1) : chunk;
      const int tiles_here = min(a.tpc, a.ntiles - chunk * a.tpc);
      const int nitems = tiles_here * a.ncls * npiece;
      if (fast) {
        int opos_t[4], prow_t[4];
        float imw_t[4];
        bool valid_t[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int p = a.row_first + (chunk * a.tpc + t) * 128 + warp * 32 + lane;
          const int h2 = p % a.H[2];
          const int h01 = p / a.H[2];
          const int h1 = h01 % a.H[1], h0 = h01 / a.H[1];
          const int r0 = org[0] + h0 - a.lo[0], r1 = org[1] + h1 - a.lo[1], r2 = org[2] + h2 - a.lo[2];
          valid_t[t] = t < tiles_here && h0 >= a.lo[0] && h0 < a.lo[0] + a.t[0] && h1 >= a.lo[1] &&
                       h1 < a.lo[1] + a.t[1] && h2 >= a.lo[2] && h2 < a.lo[2] + a.t[2] && r0 < a.rd[0] &&
                       r1 < a.rd[1] && r2 < a.rd[2];
          opos_t[t] = (r0 * a.od[1] + r1) * a.od[2] + r2;
          prow_t[t] = p;
          imw_t[t] = 1.f;
          if (a.pl_weighted && valid_t[t])
            imw_t[t] = fmaxf(__fmul_rn(__fmul_rn(a.imap0[r0], a.imap1[r1]), a.imap2[r2]), a.imap_floor);
        }
        const int gcg = (coblk * N) >> 3;
        const bool segA = gcg < a.segA_cg;
        uint4 nres0 = make_uint4(0, 0, 0, 0), nres1 = nres0;
        auto fast_res = [&](int t) {
          if (valid_t[t] && segA) {
            nres0 = __ldg(reinterpret_cast<const uint4*>(a.res + (((long long)n * a.cgA + gcg) * ovox + opos_t[t]) * 8));
            nres1 = __ldg(reinterpret_cast<const uint4*>(a.res + (((long long)n * a.cgA + gcg + 1) * ovox + opos_t[t]) * 8));
          }
        };
        if (a.res_mode == 1) fast_res(0);
        ok = mbar_wait(TFULL(buf), (uint32_t)use & 1u, a.error_flag, 4);
        if (!ok) break;
        tc_fence_after();
        if (warp == 0 && chunk < 4) SGM_TRACE(8 + chunk);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          if (t >= tiles_here) break;
          uint32_t raw[16];
          tc_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(buf * a.cols_per_buf + t * N), raw);
          const uint4 cres0 = nres0, cres1 = nres1;
          if (a.res_mode == 1 && t + 1 < tiles_here) fast_res(t + 1);
          if (!valid_t[t]) continue;
          float v[16];
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            float x = __uint_as_float(raw[c]) + bias_r[c];
            if (segA && a.actA) x = prelu(x, a.alphaA);
            v[c] = x;
          }
          if (segA) {
            if (a.res_mode == 1) {
              float r[8];
              unpack8(cres0, r);
#pragma unroll
              for (int c = 0; c < 8; ++c) v[c] += r[c];
              unpack8(cres1, r);
#pragma unroll
              for (int c = 0; c < 8; ++c) v[8 + c] += r[c];
            } else if (a.res_mode == 2) {
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                float r[8];
                unpack8(*reinterpret_cast<const uint4*>(a_smem + ((size_t)(gcg + h) * a.P + prow_t[t]) * 16), r);
#pragma unroll
                for (int c = 0; c < 8; ++c) v[h * 8 + c] += r[c];
              }
            }
            if (a.out_kind == OUT_CG8) {
#pragma unroll
              for (int h = 0; h < 2; ++h)
                if (gcg + h < a.cgA)
                  *reinterpret_cast<uint4*>(a.outA + (((long long)n * a.cgA + gcg + h) * ovox + opos_t[t]) * 8) =
                      pack8(v + h * 8);
            } else {  // OUT_PLANAR (fp32 logits, optionally importance-weighted)
              float* dst = a.pl_out + (long long)n * a.pl_nstride + opos_t[t];
#pragma unroll
              for (int c = 0; c < 16; ++c)
                if (c < a.c_real) __stcs(dst + c * a.pl_cstride, a.pl_weighted ? __fmul_rn(v[c], imw_t[t]) : v[c]);
            }
          } else {
            const int bcg = gcg - a.segA_cg;
#pragma unroll
            for (int h = 0; h < 2; ++h)
              if (bcg + h < a.cgB)
                *reinterpret_cast<uint4*>(a.outB + (((long long)n * a.cgB + bcg + h) * ovox + opos_t[t]) * 8) =
                    pack8(v + h * 8);
          }
        }
        tc_fence_before();
        mbar_arrive(TEMPTY(buf));
        if (warp == 0 && chunk < 4) SGM_TRACE(12 + chunk);
        continue;
      }
      // ---- generic path: any N / class count.  Geometry is per tile (hoisted out of the class / piece
      // loops); the global residual of the next 16-channel piece is prefetched while one is processed.
      (void)nitems;
      ok = mbar_wait(TFULL(buf), (uint32_t)use & 1u, a.error_flag, 4);
      if (!ok) break;
      tc_fence_after();
      if (warp == 0 && chunk < 4) SGM_TRACE(8 + chunk);  // accumulators of the chunk complete
      for (int t = 0; t < tiles_here; ++t) {
        const int p = a.row_first + (chunk * a.tpc + t) * 128 + warp * 32 + lane;
        const int h2 = p % a.H[2];
        const int h01 = p / a.H[2];
        const int h1 = h01 % a.H[1], h0 = h01 / a.H[1];
        const int r0 = org[0] + h0 - a.lo[0], r1 = org[1] + h1 - a.lo[1], r2 = org[2] + h2 - a.lo[2];
        const bool valid = h0 >= a.lo[0] && h0 < a.lo[0] + a.t[0] && h1 >= a.lo[1] && h1 < a.lo[1] + a.t[1] &&
                           h2 >= a.lo[2] && h2 < a.lo[2] + a.t[2] && r0 < a.rd[0] && r1 < a.rd[1] && r2 < a.rd[2];
        for (int cls = 0; cls < a.ncls; ++cls) {
          int o0 = r0, o1 = r1, o2 = r2;
          if (a.mode == MODE_T2) {
            int bits = cls;
            o2 = 2 * r2 + (bits & 1);
            bits >>= 1;
            o1 = 2 * r1 + (bits & 1);
            bits >>= 1;
            o0 = a.par[0] == 2 ? 2 * r0 + (bits & 1) : r0;
          }
          const long long opos = ((long long)o0 * a.od[1] + o1) * a.od[2] + o2;
          const uint32_t tcol0 = tmem_base + ((uint32_t)(warp * 32) << 16) +
                                 (uint32_t)(buf * a.cols_per_buf + (t * a.ncls + cls) * N);
          uint4 nres0 = make_uint4(0, 0, 0, 0), nres1 = nres0;
          auto res_fetch = [&](int piece) {
            const int g = (coblk * N + piece * 16) >> 3;
            if (valid && g < a.segA_cg) {
              nres0 = __ldg(reinterpret_cast<const uint4*>(a.res + (((long long)n * a.cgA + g) * ovox + opos) * 8));
              nres1 = __ldg(reinterpret_cast<const uint4*>(a.res + (((long long)n * a.cgA + g + 1) * ovox + opos) * 8));
            }
          };
          if (a.res_mode == 1) res_fetch(0);
          for (int piece = 0; piece < npiece; ++piece) {
            uint32_t raw[16];
            tc_ld16(tcol0 + piece * 16, raw);  // warp-collective: every lane executes it
            const uint4 cres0 = nres0, cres1 = nres1;
            if (a.res_mode == 1 && piece + 1 < npiece) res_fetch(piece + 1);
            if (!valid) continue;
            const int cbase = coblk * N + piece * 16;  // fused output channel of raw[0]
            const int gcg = cbase >> 3;
            const bool segA = gcg < a.segA_cg;
            float v[16];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.bias + cbase) + q);
              v[4 * q + 0] = __uint_as_float(raw[4 * q + 0]) + b4.x;
              v[4 * q + 1] = __uint_as_float(raw[4 * q + 1]) + b4.y;
              v[4 * q + 2] = __uint_as_float(raw[4 * q + 2]) + b4.z;
              v[4 * q + 3] = __uint_as_float(raw[4 * q + 3]) + b4.w;
            }
            if (segA && a.actA) {
#pragma unroll
              for (int c = 0; c < 16; ++c) v[c] = prelu(v[c], a.alphaA);
            }
            if (segA) {
              if (a.res_mode == 1) {
                float r[8];
                unpack8(cres0, r);
#pragma unroll
                for (int c = 0; c < 8; ++c) v[c] += r[c];
                unpack8(cres1, r);
#pragma unroll
                for (int c = 0; c < 8; ++c) v[8 + c] += r[c];
              } else if (a.res_mode == 2) {  // identity residual: centre of the brick, channel group = gcg
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                  float r[8];
                  unpack8(*reinterpret_cast<const uint4*>(a_smem + ((size_t)(gcg + h) * a.P + p) * 16), r);
#pragma unroll
                  for (int c = 0; c < 8; ++c) v[h * 8 + c] += r[c];
                }
              }
              if (a.out_kind == OUT_CG8) {
#pragma unroll
                for (int h = 0; h < 2; ++h)
                  if (gcg + h < a.cgA)
                    *reinterpret_cast<uint4*>(a.outA + (((long long)n * a.cgA + gcg + h) * ovox + opos) * 8) =
                        pack8(v + h * 8);
              } else if (a.out_kind == OUT_BLEND) {
                const int g0 = a.wo[0] + o0;
                if (g0 >= 0 && g0 < a.ad0) {
                  const long long off = ((long long)g0 * a.ad1 + (a.wo[1] + o1)) * a.ad2 + (a.wo[2] + o2);
                  const float imw = fmaxf(__fmul_rn(__fmul_rn(a.imap0[o0], a.imap1[o1]), a.imap2[o2]), a.imap_floor);
                  float oldv[16];
#pragma unroll
                  for (int c = 0; c < 16; ++c)
                    if (cbase + c < a.c_real) oldv[c] = __ldcg(a.pl_out + (cbase + c) * a.pl_cstride + off);
#pragma unroll
                  for (int c = 0; c < 16; ++c)  // seg *= w; out += seg (two roundings, as MONAI)
                    if (cbase + c < a.c_real)
                      __stcg(a.pl_out + (cbase + c) * a.pl_cstride + off, __fadd_rn(oldv[c], __fmul_rn(v[c], imw)));
                }
              } else {  // OUT_PLANAR: fp32 logits [n][C][od], optionally pre-multiplied by the importance map
                float imw = 1.f;
                if (a.pl_weighted)
                  imw = fmaxf(__fmul_rn(__fmul_rn(a.imap0[o0], a.imap1[o1]), a.imap2[o2]), a.imap_floor);
                const long long off = (long long)n * a.pl_nstride + opos;
#pragma unroll
                for (int c = 0; c < 16; ++c)
                  if (cbase + c < a.c_real)
                    __stcs(a.pl_out + (cbase + c) * a.pl_cstride + off, a.pl_weighted ? __fmul_rn(v[c], imw) : v[c]);
              }
            } else {
              const int bcg = gcg - a.segA_cg;
#pragma unroll
              for (int h = 0; h < 2; ++h)
                if (bcg + h < a.cgB)
                  *reinterpret_cast<uint4*>(a.outB + (((long long)n * a.cgB + bcg + h) * ovox + opos) * 8) =
                      pack8(v + h * 8);
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(TEMPTY(buf));
      if (warp == 0 && chunk < 4) SGM_TRACE(12 + chunk);  // epilogue of the chunk done
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) SGM_TRACE(16);
  if (warp == 5) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)a.tmem_cols)
                 : "memory");
  }
}

// ------------------------------------------------------------------------------------------- host
inline uint16_t f2bf(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7f800000u) != 0x7f800000u) u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}

inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

}  // namespace

bool tc_supported(const sgm_conv_desc& d) {
  if (d.kind == SGM_KIND_IDENTITY) return false;
  return (d.kernel == 3 || d.kernel == 1) && (d.stride == 1 || d.stride == 2);
}

struct PlanEntry;
static void free_plans(void* p);

void tc_free(TcConv* c) {
  if (!c) return;
  free_plans(c->plan_cache);
  ps_free(c);
  pst_free(c);
  rs_free(c);
  cs_free(c);
  if (c->w) cudaFree(c->w);
  if (c->bias) cudaFree(c->bias);
  delete c;
}

// K-block descriptor tables depend only on (mode, kernel extents, channel pairs): identical tables of
// different convolutions / networks share one slice of the constant bank.
struct BlkSlice {
  std::vector<uint32_t> words;
  int off;
};
static std::vector<BlkSlice> g_blk_slices_dev[64];  // the __constant__ bank is per device
static int g_blk_used_dev[64] = {0};

static int blk_offset(const TcConv& c) {
  int dev = 0;
  cudaGetDevice(&dev);
  GlobalLock lock;
  std::vector<BlkSlice>& g_blk_slices = g_blk_slices_dev[dev & 63];
  int& g_blk_used = g_blk_used_dev[dev & 63];
  std::vector<uint32_t> d(c.blocks.size());
  for (size_t i = 0; i < d.size(); ++i) {
    const KBlock& b = c.blocks[i];
    d[i] = (uint32_t)(b.shift[0] + 1) | ((uint32_t)(b.shift[1] + 1) << 2) | ((uint32_t)(b.shift[2] + 1) << 4) |
           ((uint32_t)b.slab << 6) | ((uint32_t)b.cls << 9) | ((uint32_t)b.first << 12) | ((uint32_t)b.cgpair << 16);
  }
  for (auto& sl : g_blk_slices)
    if (sl.words == d) return sl.off;
  if (g_blk_used + (int)d.size() > kBlkConst) return -1;
  const int off = g_blk_used;
  if (cudaMemcpyToSymbol(c_blk, d.data(), d.size() * 4, (size_t)off * 4, cudaMemcpyHostToDevice) != cudaSuccess) return -1;
  g_blk_used += (int)d.size();
  g_blk_slices.push_back({d, off});
  return off;
}

// Weight element of the (possibly flipped / transposed) conv: W(co, ci, tap)
static inline float wget(const sgm_conv_desc& d, int co, int ci, int tap, int ntaps) {
  if (d.kind == SGM_KIND_CONV_TRANSPOSE) {
    const int t = d.stride == 1 ? (ntaps - 1 - tap) : tap;  // stride-1 transposed conv == flipped conv
    return d.weight[((size_t)ci * d.cout + co) * ntaps + t];
  }
  return d.weight[((size_t)co * d.cin + ci) * ntaps + tap];
}

int tc_pack(const sgm_conv_desc* m, const sgm_conv_desc* second, int spatial_dims, TcConv** out) {
  *out = nullptr;
  SGM_REQUIRE(m && tc_supported(*m), SGM_ERR_UNSUPPORTED, "conv not supported by the tcgen05 family");
  TcConv* c = new TcConv();
  c->plan_cache = new std::vector<PlanEntry>();
  const bool tr2 = m->kind == SGM_KIND_CONV_TRANSPOSE && m->stride == 2;
  c->mode = tr2 ? MODE_T2 : (m->stride == 2 ? MODE_S2 : MODE_S1);
  c->flat0 = spatial_dims == 2;
  c->cin = m->cin;
  c->cgin = round_up(m->cin, 16) / 8;
  for (int a = 0; a < 3; ++a) c->k[a] = (c->flat0 && a == 0) ? 1 : m->kernel;
  const int ntaps = c->k[0] * c->k[1] * c->k[2];
  const int nA = round_up(m->cout, 16);
  int nB = 0;
  if (second) {
    if (second->kind != SGM_KIND_CONV || second->kernel != m->kernel || second->stride != m->stride ||
        second->cin != m->cin || m->kind != SGM_KIND_CONV) {
      delete c;
      set_error("tc_pack: second conv cannot be fused (different geometry)");
      return SGM_ERR_INVALID;
    }
    nB = round_up(second->cout, 16);
  }
  c->ntot = nA + nB;
  c->segA_cg = nA / 8;
  c->actA = m->has_act;
  c->alphaA = m->alpha;
  c->c_real = m->cout;
  c->ncls = tr2 ? (c->flat0 ? 4 : 8) : 1;
  // output channels per CTA: multiple of 16 dividing ntot; TMEM: ncls * N <= 256 columns
  const int cap = tr2 ? std::max(16, 256 / c->ncls) : 128;
  // deep stride-1 layers run one CTA per (window, column block): 64 columns keep all SMs busy at 6^3
  const bool deep = c->cgin >= 16 && c->cgin % 4 == 0 && !c->flat0 && !getenv("SGM_NO_ACHUNK");
  const bool deep_s1 = deep && !tr2 && m->stride == 1;
  // deep transposed layers: 8 classes x 16 columns x 3 tiles of the 7^3 brick = 384 TMEM columns, all resident
  const bool deep_t2 = deep && tr2 && c->cgin >= 32 && !getenv("SGM_NO_ACHUNK_T2");
  int N = 16;
  for (int cand = 16; cand <= std::min(deep_s1 ? 64 : (deep_t2 ? 16 : cap), c->ntot); cand += 16)
    if (c->ntot % cand == 0) N = cand;
  c->ncta = N;
  c->ncoblk = c->ntot / N;

  // ---- K blocks
  const int ncgp = c->cgin / 2;
  c->kmajor = deep_s1 || deep_t2;
  if (deep_t2) {
    // channel-major K blocks of the class-folded transposed conv: (cgpair, input shift), all 8 parity classes of
    // the CTA's 16 output channels in one MMA (N = 128, zero weights where a class does not read the shift)
    c->tfold = 1;
    for (int p = 0; p < ncgp; ++p)
      for (int sh0 = 0; sh0 < 2; ++sh0)
        for (int sh1 = 0; sh1 < 2; ++sh1)
          for (int sh2 = 0; sh2 < 2; ++sh2) {
            KBlock b{0, p, {sh0, sh1, sh2}, 0, c->blocks.empty() ? 1 : 0};
            c->blocks.push_back(b);
          }
  } else if (c->kmajor) {
    // deep layers (>= 128 input channels): channel-major K blocks, the brick is streamed chunk by chunk (tc_plan)
    for (int p = 0; p < ncgp; ++p)
      for (int k0 = 0; k0 < c->k[0]; ++k0)
        for (int k1 = 0; k1 < c->k[1]; ++k1)
          for (int k2 = 0; k2 < c->k[2]; ++k2) {
            KBlock b{0, p, {k0 - c->k[0] / 2, k1 - c->k[1] / 2, k2 - c->k[2] / 2}, 0, c->blocks.empty() ? 1 : 0};
            c->blocks.push_back(b);
          }
  } else if (c->mode == MODE_S1) {
    for (int k0 = 0; k0 < c->k[0]; ++k0)
      for (int k1 = 0; k1 < c->k[1]; ++k1)
        for (int k2 = 0; k2 < c->k[2]; ++k2)
          for (int p = 0; p < ncgp; ++p) {
            KBlock b{0, p, {k0 - c->k[0] / 2, k1 - c->k[1] / 2, k2 - c->k[2] / 2}, 0, c->blocks.empty() ? 1 : 0};
            c->blocks.push_back(b);
          }
  } else if (c->mode == MODE_S2) {
    auto pe = [](int K, int k, int& r, int& e) {  // i = 2o + k - 1 = 2(o+e) + r
      if (K == 1) { r = 0; e = 0; return; }
      if (k == 0) { r = 1; e = -1; } else if (k == 1) { r = 0; e = 0; } else { r = 1; e = 0; }
    };
    for (int k0 = 0; k0 < c->k[0]; ++k0)
      for (int k1 = 0; k1 < c->k[1]; ++k1)
        for (int k2 = 0; k2 < c->k[2]; ++k2) {
          int r0, e0, r1, e1, r2, e2;
          pe(c->k[0], k0, r0, e0), pe(c->k[1], k1, r1, e1), pe(c->k[2], k2, r2, e2);
          const int slab = c->flat0 ? (r1 * 2 + r2) : (r0 * 4 + r1 * 2 + r2);
          for (int p = 0; p < ncgp; ++p) {
            KBlock b{slab, p, {e0, e1, e2}, 0, c->blocks.empty() ? 1 : 0};
            c->blocks.push_back(b);
          }
        }
  } else if (c->ncls * N <= 128 && !getenv("SGM_NO_TFOLD")) {
    // Transposed conv with few output channels: the parity classes that read the same input shift share one MMA
    // (N = ncls * Cout columns, zero weights where a class does not use the shift) -- 8 instead of 27 MMAs per 16
    // input channels; an M=128 K=16 MMA costs ~34 + 0.36 N clk (tests/ubench_mma.cu), so N = 128 is only 2x N = 16.
    c->tfold = 1;
    const int np0 = c->flat0 ? 1 : 2;
    for (int sh0 = 0; sh0 < np0; ++sh0)
      for (int sh1 = 0; sh1 < 2; ++sh1)
        for (int sh2 = 0; sh2 < 2; ++sh2)
          for (int p = 0; p < ncgp; ++p) {
            KBlock b{0, p, {sh0, sh1, sh2}, 0, c->blocks.empty() ? 1 : 0};
            c->blocks.push_back(b);
          }
  } else {
    // o = 2j + p reads in[j + sh] * W[k]:  p=0: (sh 0, k 1);  p=1: (sh 0, k 2), (sh 1, k 0)
    const int np0 = c->flat0 ? 1 : 2;
    for (int p0 = 0; p0 < np0; ++p0)
      for (int p1 = 0; p1 < 2; ++p1)
        for (int p2 = 0; p2 < 2; ++p2) {
          const int cls = c->flat0 ? (p1 * 2 + p2) : (p0 * 4 + p1 * 2 + p2);
          bool first = true;
          for (int sh0 = 0; sh0 <= (c->flat0 ? 0 : p0); ++sh0)
            for (int sh1 = 0; sh1 <= p1; ++sh1)
              for (int sh2 = 0; sh2 <= p2; ++sh2)
                for (int p = 0; p < ncgp; ++p) {
                  KBlock b{0, p, {sh0, sh1, sh2}, cls, first ? 1 : 0};
                  first = false;
                  c->blocks.push_back(b);
                }
        }
  }
  const int nblk = (int)c->blocks.size();

  // ---- weights: [coblk][blk][kchunk 2][N][8] bf16 (class-folded transposed conv: [blk][kchunk 2][cls][N][8])
  const int NB = c->tfold ? c->ncls * N : N;
  c->mma_n = NB;
  std::vector<uint16_t> w((size_t)c->ncoblk * nblk * 2 * NB * 8, 0);
  std::vector<float> bias(c->ntot, 0.f);
  for (int co = 0; co < m->cout; ++co) bias[co] = m->bias[co];
  if (second)
    for (int co = 0; co < second->cout; ++co) bias[nA + co] = second->bias[co];
  auto tap_of = [&](const KBlock& b) -> int {
    int kk[3];
    for (int a = 0; a < 3; ++a) {
      if (c->k[a] == 1) { kk[a] = 0; continue; }
      if (c->mode == MODE_S1) kk[a] = b.shift[a] + 1;
      else if (c->mode == MODE_S2) {
        // recover k from (r, e): slab bit r, shift e
        int bit;
        if (c->flat0) bit = a == 1 ? (b.slab >> 1) & 1 : (b.slab & 1);
        else bit = (b.slab >> (2 - a)) & 1;
        kk[a] = bit == 0 ? 1 : (b.shift[a] == -1 ? 0 : 2);
      } else {
        int pbit;
        if (c->flat0) pbit = a == 1 ? (b.cls >> 1) & 1 : (b.cls & 1);
        else pbit = (b.cls >> (2 - a)) & 1;
        kk[a] = pbit == 0 ? 1 : (b.shift[a] == 0 ? 2 : 0);
      }
    }
    return (kk[0] * c->k[1] + kk[1]) * c->k[2] + kk[2];
  };
  for (int cb = 0; cb < c->ncoblk && c->tfold; ++cb) {
    for (int bi = 0; bi < nblk; ++bi) {
      const KBlock& b = c->blocks[bi];
      for (int cls = 0; cls < c->ncls; ++cls) {
        // tap of this class for the block's input shift; a class with parity bit 0 only reads shift 0
        int kk[3];
        bool used = true;
        for (int a = 0; a < 3; ++a) {
          if (c->k[a] == 1) { kk[a] = 0; continue; }
          const int pbit = c->flat0 ? (a == 1 ? (cls >> 1) & 1 : (cls & 1)) : (cls >> (2 - a)) & 1;
          if (pbit == 0) { kk[a] = 1; used = used && b.shift[a] == 0; }
          else kk[a] = b.shift[a] == 0 ? 2 : 0;
        }
        if (!used) continue;
        const int tap = (kk[0] * c->k[1] + kk[1]) * c->k[2] + kk[2];
        for (int kc = 0; kc < 2; ++kc)
          for (int nn = 0; nn < N; ++nn) {
            const int co = cb * N + nn;
            if (co >= m->cout) continue;
            for (int k8 = 0; k8 < 8; ++k8) {
              const int ci = (b.cgpair * 2 + kc) * 8 + k8;
              if (ci >= m->cin) continue;
              w[((((size_t)cb * nblk + bi) * 2 + kc) * NB + (cls * N + nn)) * 8 + k8] = f2bf(wget(*m, co, ci, tap, ntaps));
            }
          }
      }
    }
  }
  for (int cb = 0; cb < c->ncoblk && !c->tfold; ++cb)
    for (int bi = 0; bi < nblk; ++bi) {
      const KBlock& b = c->blocks[bi];
      const int tap = tap_of(b);
      for (int kc = 0; kc < 2; ++kc)
        for (int nn = 0; nn < N; ++nn) {
          const int fco = cb * N + nn;  // fused output channel
          const sgm_conv_desc* src = fco < nA ? m : second;
          const int co = fco < nA ? fco : fco - nA;
          if (!src || co >= src->cout) continue;
          for (int k8 = 0; k8 < 8; ++k8) {
            const int ci = (b.cgpair * 2 + kc) * 8 + k8;
            if (ci >= m->cin) continue;
            w[((((size_t)cb * nblk + bi) * 2 + kc) * N + nn) * 8 + k8] = f2bf(wget(*src, co, ci, tap, ntaps));
          }
        }
    }
  if (cudaMalloc(&c->w, w.size() * 2) != cudaSuccess || cudaMalloc(&c->bias, bias.size() * 4) != cudaSuccess) {
    set_error("tc_pack: cudaMalloc failed: %s", cudaGetErrorString(cudaGetLastError()));
    tc_free(c);
    return SGM_ERR_CUDA;
  }
  cudaMemcpy(c->w, w.data(), w.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(c->bias, bias.data(), bias.size() * 4, cudaMemcpyHostToDevice);
  c->blk_off = blk_offset(*c);
  if (c->blk_off < 0) {
    set_error("tc_pack: K-block descriptor bank exhausted / upload failed");
    tc_free(c);
    return SGM_ERR_CUDA;
  }
  {
    int rc = cs_pack(m, second, c);
    if (rc) {
      tc_free(c);
      return rc;
    }
  }
  if (!second) {
    int rc = ps_pack(*m, c);
    if (!rc) rc = pst_pack(*m, c);
    if (!rc) rc = rs_pack(*m, c);
    if (rc) {
      tc_free(c);
      return rc;
    }
  }
  *out = c;
  return SGM_OK;
}

// ---- TMA tensor maps for the halo bricks.  A CG8 tensor [n*cg][D0][D1][D2][8 bf16] is a rank-4 tiled
// tensor {D2*8, D1, D0, n*cg}; one box = one channel group of the brick {H2*8, H1, H0, 1}.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    if (getenv("SGM_NO_TMA")) return nullptr;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

bool tma_available() { return encode_fn() != nullptr; }

struct MapKey {
  const void* ptr;
  int ncg, d[3], h[3], par[3];
};
struct MapEntry {
  MapKey key;
  CUtensorMap map;
};
static std::vector<MapEntry> g_maps;

int make_brick_map(CUtensorMap* out, const void* ptr, int ncg, const int d[3], const int H[3], const int* par) {
  GlobalLock lock;
  MapKey key{ptr, ncg, {d[0], d[1], d[2]}, {H[0], H[1], H[2]}, {par ? par[0] : 0, par ? par[1] : 0, par ? par[2] : 0}};
  for (auto& e : g_maps)
    if (memcmp(&e.key, &key, sizeof(key)) == 0) {
      *out = e.map;
      return SGM_OK;
    }
  if (par) {
    // stride-2 parity slab: rank 5 {8 ch, D2, D1, D0, n*cg}; the box walks 2*H voxels with element stride 2
    const cuuint64_t gdim5[5] = {8, (cuuint64_t)d[2], (cuuint64_t)d[1], (cuuint64_t)d[0], (cuuint64_t)ncg};
    const cuuint64_t gstr5[4] = {16, (cuuint64_t)d[2] * 16, (cuuint64_t)d[1] * d[2] * 16,
                                 (cuuint64_t)d[0] * d[1] * d[2] * 16};
    const cuuint32_t box5[5] = {8, (cuuint32_t)(H[2] * par[2]), (cuuint32_t)(H[1] * par[1]),
                                (cuuint32_t)(H[0] * par[0]), 1};
    const cuuint32_t estr5[5] = {1, (cuuint32_t)par[2], (cuuint32_t)par[1], (cuuint32_t)par[0], 1};
    CUresult r5 = encode_fn()(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(ptr), gdim5, gstr5, box5,
                              estr5, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r5 != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled (strided) failed (%d) for dims (%d,%d,%d), box (%d,%d,%d)", (int)r5, d[0], d[1],
                d[2], H[0], H[1], H[2]);
      return SGM_ERR_CUDA;
    }
    if (g_maps.size() > 4096) g_maps.clear();
    g_maps.push_back({key, *out});
    return SGM_OK;
  }
  const cuuint64_t gdim[4] = {(cuuint64_t)d[2] * 8, (cuuint64_t)d[1], (cuuint64_t)d[0], (cuuint64_t)ncg};
  const cuuint64_t gstr[3] = {(cuuint64_t)d[2] * 16, (cuuint64_t)d[1] * d[2] * 16, (cuuint64_t)d[0] * d[1] * d[2] * 16};
  const cuuint32_t box[4] = {(cuuint32_t)H[2] * 8, (cuuint32_t)H[1], (cuuint32_t)H[0], 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = encode_fn()(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), gdim, gstr, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) for dims (%d,%d,%d) x %d groups, box (%d,%d,%d)", (int)r, d[0],
              d[1], d[2], ncg, H[0], H[1], H[2]);
    return SGM_ERR_CUDA;
  }
  if (g_maps.size() > 4096) g_maps.clear();
  g_maps.push_back({key, *out});
  return SGM_OK;
}

// Geometry of one launch (tile shape search, TMEM / weight-ring sizing): depends only on the conv and the
// tensor extents, so it is computed once per (conv, dims, batch) and cached.
static int tc_plan(const TcConv& c, const TcIO& io, KArgs& a, int& smem_bytes_out, int ctas) {
  memset(&a, 0, sizeof(a));
  SGM_REQUIRE(io.cg0 + io.cg1 == c.cgin, SGM_ERR_INVALID, "tc_launch: input channel groups %d+%d != %d", io.cg0,
              io.cg1, c.cgin);
  a.cg0 = io.cg0, a.cg1 = io.cg1, a.cgin = c.cgin;
  a.mode = c.mode;
  for (int i = 0; i < 3; ++i) {
    a.id[i] = io.id[i], a.od[i] = io.od[i];
    a.rd[i] = c.mode == MODE_T2 ? io.id[i] : io.od[i];
  }
  const int N = c.ncta, nblk = (int)c.blocks.size();
  a.N = N, a.nblk = nblk, a.ncls = c.ncls;
  const int NB = c.mma_n;
  a.mmaN = NB;
  // weight ring
  const bool three = ctas == 3 && !c.kmajor;
  a.G = std::max(1, std::min(nblk, (three ? 8192 : 16384) / (NB * 32)));
  a.ngroups = ceil_div(nblk, a.G);
  a.nstages = std::min(three ? 3 : (a.G * NB * 32 <= 8192 ? 6 : 4), a.ngroups);
  a.resident = a.ngroups <= a.nstages;
  a.w_stage_bytes = round_up(a.G * NB * 32, 128);
  // TMEM
  int cols_tile = c.ncls * N;
  a.nbuf = cols_tile <= 128 ? 2 : 1;
  a.tpc = std::max(1, (a.nbuf == 2 ? 128 : 256) / cols_tile);
  if (c.ncls == 1 && N == 16) a.tpc = 4;  // epilogue fast path: 4 tiles per TMEM buffer
  SGM_REQUIRE(cols_tile <= 256, SGM_ERR_UNSUPPORTED, "tc_launch: %d TMEM columns per tile", cols_tile);

  // ---- per-axis geometry of the brick
  int pad[3], addH[3];
  for (int i = 0; i < 3; ++i) {
    const bool flat = c.k[i] == 1 && (c.flat0 && i == 0);
    pad[i] = c.k[i] / 2;
    if (c.mode == MODE_S1) {
      a.par[i] = 1, a.lo[i] = pad[i], addH[i] = 2 * pad[i];
      a.ibase_mul[i] = 1, a.ioff[i] = -pad[i], a.imul[i] = 1;
    } else if (c.mode == MODE_S2) {
      if (flat || c.k[i] == 1) {
        a.par[i] = 1, a.lo[i] = 0, addH[i] = 0, a.ibase_mul[i] = 1, a.ioff[i] = 0, a.imul[i] = 1;
      } else {
        a.par[i] = 2, a.lo[i] = 1, addH[i] = 1, a.ibase_mul[i] = 2, a.ioff[i] = -2, a.imul[i] = 2;
      }
    } else {
      if (flat) {
        a.par[i] = 1, a.lo[i] = 0, addH[i] = 0;
      } else {
        a.par[i] = 2, a.lo[i] = 0, addH[i] = 1;  // par==2 marks "output = 2*row + class bit"
      }
      a.ibase_mul[i] = 1, a.ioff[i] = 0, a.imul[i] = 1;
    }
  }
  a.nslab = 1;
  if (c.mode == MODE_S2)
    for (int i = 0; i < 3; ++i) a.nslab *= a.par[i];

  // ---- tile shape search (cost model: MMA issue cycles + brick bytes, times waves)
  static const int cand[] = {1, 2, 3, 4, 6, 8, 12, 16, 24, 32, 48, 64, 96};
  double best = 1e30;
  int bt[3] = {0, 0, 0};
  const int fixed_bytes = a.nstages * a.w_stage_bytes + (2 * a.nstages + 14) * 8 + nblk * 8 + 16 + 256;
  a.use_tma = tma_available() && !(c.mode == MODE_S2 && getenv("SGM_NO_TMA_S2"));
  if (c.kmajor) {
    // ---- chunked-A plan: one CTA per SM, all tiles of the brick accumulate in TMEM at once
    SGM_REQUIRE(a.use_tma, SGM_ERR_UNSUPPORTED, "deep layers need TMA (cuTensorMapEncodeTiled unavailable)");
    a.achunk = 4;  // 32 input channels per ring stage
    a.nkc = c.cgin / a.achunk;
    a.astages = std::min(3, a.nkc);
    const int budget = 200 * 1024;
    for (int c0 : cand)
      for (int c1 : cand)
        for (int c2 : cand) {
          const int t[3] = {std::min(c0, a.rd[0]), std::min(c1, a.rd[1]), std::min(c2, a.rd[2])};
          const int H[3] = {t[0] + addH[0], t[1] + addH[1], t[2] + addH[2]};
          const int P = round_up(H[0] * H[1] * H[2], 8);
          if (P > 4000 || H[2] * 8 > 256 || H[1] > 256 || H[0] > 256) continue;
          const int rf = (a.lo[0] * H[1] + a.lo[1]) * H[2] + a.lo[2];
          const int rl = ((a.lo[0] + t[0] - 1) * H[1] + a.lo[1] + t[1] - 1) * H[2] + a.lo[2] + t[2] - 1;
          const int ntl = ceil_div(rl - rf + 1, 128);
          if (ntl * cols_tile > 512) continue;  // all tiles (x parity classes) stay resident in TMEM
          const int su = round_up(a.achunk * P + 128 + 2 * (H[1] * H[2] + H[2] + 2), 8);
          if ((long long)a.astages * su * 16 + fixed_bytes > budget) continue;
          const long long nct = (long long)ceil_div(a.rd[0], t[0]) * ceil_div(a.rd[1], t[1]) * ceil_div(a.rd[2], t[2]) *
                                c.ncoblk * io.n;
          const double mma = (double)ntl * nblk * (34.0 + 0.36 * NB);
          const double cta = mma + 6000.0;
          const double waves = (double)((nct + 147) / 148);
          const double cost = waves * cta;
          if (cost < best) best = cost, bt[0] = t[0], bt[1] = t[1], bt[2] = t[2];
        }
    SGM_REQUIRE(bt[0] > 0, SGM_ERR_UNSUPPORTED, "tc_launch: no chunked tile shape fits (cgin=%d, N=%d)", c.cgin, N);
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
    a.tpc = a.ntiles, a.nchunks = 1, a.nbuf = 1;
    a.cols_per_buf = a.ntiles * cols_tile;
    int pw = 32;
    while (pw < a.cols_per_buf) pw <<= 1;
    a.tmem_cols = pw;
    a.astage_units = round_up(a.achunk * a.P + 128 + 2 * (a.H[1] * a.H[2] + a.H[2] + 2), 8);
    a.a_units = a.astages * a.astage_units;
    smem_bytes_out = a.a_units * 16 + fixed_bytes;
    return SGM_OK;
  }
  for (int c0 : cand)
    for (int c1 : cand)
      for (int c2 : cand) {
        const int t[3] = {std::min(c0, a.rd[0]), std::min(c1, a.rd[1]), std::min(c2, a.rd[2])};
        const int H[3] = {t[0] + addH[0], t[1] + addH[1], t[2] + addH[2]};
        const int P = round_up(H[0] * H[1] * H[2], 8);
        if (P > 16383) continue;
        if (a.use_tma && c.mode != MODE_S2 && (H[2] * 8 > 256 || H[1] > 256 || H[0] > 256)) continue;  // TMA box limits
        if (a.use_tma && c.mode == MODE_S2 && (H[2] * 2 > 256 || H[1] * 2 > 256 || H[0] * 2 > 256)) continue;
        const int rf = (a.lo[0] * H[1] + a.lo[1]) * H[2] + a.lo[2];
        const int rl = ((a.lo[0] + t[0] - 1) * H[1] + a.lo[1] + t[1] - 1) * H[2] + a.lo[2] + t[2] - 1;
        const int ntl = ceil_div(rl - rf + 1, 128);
        const int units = a.nslab * c.cgin * P + 128 + 2 * (H[1] * H[2] + H[2] + 2);
        if ((long long)units * 16 + fixed_bytes > (ctas == 3 ? 74 * 1024 : kSmemBudget)) continue;
        const long long nct = (long long)ceil_div(a.rd[0], t[0]) * ceil_div(a.rd[1], t[1]) * ceil_div(a.rd[2], t[2]) *
                              c.ncoblk * io.n;
        const double mma = (double)ntl * nblk * (34.0 + 0.36 * NB);
        const double load = (double)a.nslab * c.cgin * P * 16 / 24.0;
        const double epi = (double)ntl * c.ncls * (N / 16) * 60.0;
        const double wload = (double)nblk * NB * 32 / 40.0;  // every CTA streams the whole filter bank from L2
        const double cta = std::max(std::max(mma, epi), wload) + load + 4000.0;
        const double waves = (double)((nct + 148 * ctas - 1) / (148 * ctas));
        const double cost = waves * cta;
        if (cost < best) best = cost, bt[0] = t[0], bt[1] = t[1], bt[2] = t[2];
      }
  SGM_REQUIRE(bt[0] > 0, SGM_ERR_UNSUPPORTED, "tc_launch: no tile shape fits shared memory (cgin=%d, N=%d)", c.cgin, N);
  for (int i = 0; i < 3; ++i) {
    a.t[i] = bt[i];
    a.H[i] = bt[i] + addH[i];
    a.nt[i] = ceil_div(a.rd[i], bt[i]);
  }
  a.P = round_up(a.H[0] * a.H[1] * a.H[2], 8);  // slab stride (16-byte units), 128-byte aligned
  a.box_bytes = a.H[0] * a.H[1] * a.H[2] * 16;
  a.row_first = (a.lo[0] * a.H[1] + a.lo[1]) * a.H[2] + a.lo[2];
  const int row_last = ((a.lo[0] + a.t[0] - 1) * a.H[1] + a.lo[1] + a.t[1] - 1) * a.H[2] + a.lo[2] + a.t[2] - 1;
  a.ntiles = ceil_div(row_last - a.row_first + 1, 128);
  a.tpc = std::min(a.tpc, a.ntiles);
  a.nchunks = ceil_div(a.ntiles, a.tpc);
  a.cols_per_buf = a.tpc * cols_tile;
  if (a.nchunks == 1) a.nbuf = 1;  // a single chunk never uses the second accumulator buffer: half the TMEM columns
  int cols = a.nbuf * a.cols_per_buf, pw = 32;
  while (pw < cols) pw <<= 1;
  a.tmem_cols = pw;
  a.a_units = a.nslab * c.cgin * a.P + 128 + 2 * (a.H[1] * a.H[2] + a.H[2] + 2);
  a.a_units = round_up(a.a_units, 8);
  const int smem_bytes = a.a_units * 16 + fixed_bytes;
  smem_bytes_out = smem_bytes;
  return SGM_OK;
}

struct PlanEntry {
  int key[8];
  KArgs args;
  int smem_bytes;
  int ctas;
};

static void free_plans(void* p) { delete reinterpret_cast<std::vector<PlanEntry>*>(p); }

int tc_launch(const TcConv& c, const TcIO& io, int* error_flag_dev, cudaStream_t st) {
  SGM_REQUIRE(io.cg0 + io.cg1 == c.cgin, SGM_ERR_INVALID, "tc_launch: input channel groups %d+%d != %d", io.cg0,
              io.cg1, c.cgin);
  if (rs_applicable(c, io)) return rs_launch(c, io, error_flag_dev, st);
  if (ps_applicable(c, io)) return ps_launch(c, io, error_flag_dev, st);
  if (pst_applicable(c, io)) return pst_launch(c, io, error_flag_dev, st);
  if (cs_applicable(c, io)) return cs_launch(c, io, error_flag_dev, st);
  const int key[8] = {io.id[0], io.id[1], io.id[2], io.od[0], io.od[1], io.od[2], io.n, io.cg0};
  auto* plans = reinterpret_cast<std::vector<PlanEntry>*>(c.plan_cache);
  const PlanEntry* pe = nullptr;
  for (auto& e : *plans)
    if (memcmp(e.key, key, sizeof(key)) == 0) pe = &e;
  if (!pe) {
    PlanEntry e;
    memcpy(e.key, key, sizeof(key));
    // three CTAs per SM where a tile's accumulators fit a third of the TMEM (<= 128 columns) and the brick plus a
    // short weight ring fit 74 KB; otherwise two with the deeper ring (measured per layer: profiles/r01c_rs_notes.md 5)
    e.ctas = target_ctas();
    int rc = e.ctas == 3 ? tc_plan(c, io, e.args, e.smem_bytes, 3) : SGM_ERR_UNSUPPORTED;
    if (rc || e.args.tmem_cols > 128 || e.args.achunk) {
      e.ctas = 2;
      rc = tc_plan(c, io, e.args, e.smem_bytes, 2);
    }
    if (rc) return rc;
    plans->push_back(e);
    pe = &plans->back();
  }
  KArgs a = pe->args;
  const int smem_bytes = pe->smem_bytes;
  const int N = c.ncta, nblk = (int)c.blocks.size();
  a.in0 = (const __nv_bfloat16*)io.in0, a.in1 = (const __nv_bfloat16*)io.in1;
  a.blk_off = c.blk_off;
  a.w = c.w, a.bias = c.bias;
  a.outA = (__nv_bfloat16*)io.outA, a.cgA = io.cgA, a.outB = (__nv_bfloat16*)io.outB, a.cgB = io.cgB;
  a.segA_cg = c.segA_cg, a.actA = c.actA, a.alphaA = c.alphaA;
  a.res = (const __nv_bfloat16*)io.res;
  a.res_mode = 0;
  if (io.res) a.res_mode = (io.res == io.in0 && c.mode == MODE_S1 && io.in1 == nullptr && io.cgA == c.cgin && !a.achunk) ? 2 : 1;
  a.pl_weighted = io.pl_weighted;
  a.out_kind = io.out_kind, a.pl_out = io.pl_out, a.pl_cstride = io.pl_cstride, a.pl_nstride = io.pl_nstride;
  a.ad0 = io.ad0, a.ad1 = io.ad1, a.ad2 = io.ad2;
  for (int i = 0; i < 3; ++i) a.wo[i] = io.wo[i];
  a.imap0 = io.imap[0], a.imap1 = io.imap[1], a.imap2 = io.imap[2], a.imap_floor = io.imap_floor;
  a.c_real = c.c_real;
  a.error_flag = error_flag_dev;

  static const bool dbg = getenv("SGM_DEBUG") != nullptr;
  if (dbg)
    fprintf(stderr,
            "[tc_launch] mode=%d N=%d nblk=%d ncls=%d cgin=%d rd=(%d,%d,%d) t=(%d,%d,%d) H=(%d,%d,%d) P=%d nslab=%d "
            "row_first=%d ntiles=%d tpc=%d nchunks=%d nbuf=%d cols/buf=%d tmem=%d G=%d ngroups=%d nstages=%d res=%d "
            "a_units=%d smem=%d tma=%d grid=(%d,%d,%d)\n",
            a.mode, N, nblk, a.ncls, a.cgin, a.rd[0], a.rd[1], a.rd[2], a.t[0], a.t[1], a.t[2], a.H[0], a.H[1], a.H[2],
            a.P, a.nslab, a.row_first, a.ntiles, a.tpc, a.nchunks, a.nbuf, a.cols_per_buf, a.tmem_cols, a.G, a.ngroups,
            a.nstages, a.resident, a.a_units, smem_bytes, a.use_tma, a.nt[0] * a.nt[1] * a.nt[2], c.ncoblk, io.n);
  static DeviceOnce attr_once;
  if (attr_once.first()) {
    SGM_CUDA_CHECK(cudaFuncSetAttribute(tc_conv_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget));
    SGM_CUDA_CHECK(cudaFuncSetAttribute(tc_conv_kernel<false, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget));
    SGM_CUDA_CHECK(cudaFuncSetAttribute(tc_conv_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
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
  if (a.use_tma) {
    int rc = make_brick_map(&tm0, io.in0, io.n * io.cg0, io.id, a.H, c.mode == MODE_S2 ? a.par : nullptr);
    if (!rc && io.in1) rc = make_brick_map(&tm1, io.in1, io.n * io.cg1, io.id, a.H);
    if (rc) return rc;
  }
  dim3 grid(a.nt[0] * a.nt[1] * a.nt[2], c.ncoblk, io.n);
  if (a.achunk) tc_conv_kernel<true><<<grid, kThreads, smem_bytes, st>>>(a, tm0, tm1);
  else if (pe->ctas == 3) tc_conv_kernel<false, 3><<<grid, kThreads, smem_bytes, st>>>(a, tm0, tm1);
  else tc_conv_kernel<false><<<grid, kThreads, smem_bytes, st>>>(a, tm0, tm1);
  SGM_CUDA_CHECK(cudaGetLastError());
  if (trace_on) {
    long long t[32];
    cudaStreamSynchronize(st);
    cudaMemcpy(t, trace_dev, sizeof(t), cudaMemcpyDeviceToHost);
    auto d = [&](int i) { return t[i] ? (double)(t[i] - t[0]) : -1.0; };
    fprintf(stderr,
            "[trace] mode=%d N=%d nblk=%d grid=%d t=(%d,%d,%d) ntiles=%d nchunks=%d | setup %.0f brick %.0f w0 %.0f | "
            "issued %.0f %.0f %.0f %.0f | accdone %.0f %.0f %.0f %.0f | epidone %.0f %.0f %.0f %.0f | end %.0f cycles\n",
            a.mode, N, nblk, grid.x * grid.y * grid.z, a.t[0], a.t[1], a.t[2], a.ntiles, a.nchunks, d(1), d(2), d(3), d(4),
            d(5), d(6), d(7), d(8), d(9), d(10), d(11), d(12), d(13), d(14), d(15), d(16));
  }
  return SGM_OK;
}

}  // namespace tc
}  // namespace sgm
