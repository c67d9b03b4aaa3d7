// Plane-sweep tcgen05 kernel for the stride-2 TRANSPOSED convolutions with few output channels (the two finest
// up-sampling layers: 64 -> 16 at 24^3 -> 48^3 and 32 -> C at 48^3 -> 96^3, 16 % of the UNet's FLOPs).
//
// out[2j + p] = sum_{s <= p} W[k(p, s)] * in[j + s]  (per axis: p = 0 reads shift 0 with tap 1; p = 1 reads shift 0
// with tap 2 and shift 1 with tap 0).  GEMM rows are INPUT voxels j; the 8 output parity classes p are folded into
// the MMA N dimension (N = 8 x 16 columns, zero weights where a class does not read a shift): 8 MMAs per 16 input
// channels instead of 27 -- an M=128 K=16 MMA costs ~34 + 0.36 N clk (tests/ubench_mma.cu), N = 128 is only twice
// N = 16.  The input is the channel concat of the skip and the sub-network tensors (two tensor maps; torch.cat is
// never materialised).
//
// Same structure as conv_ps.cu: a persistent CTA per SM sweeps a column of the window along d0, input plane by input
// plane through a TMA ring (tile x0 reads planes x0 and x0 + 1); weights stay resident in shared memory; two issuer
// warps alternate planes with the barrier polls of the next tile hidden inside the MMA burst; 16 epilogue warps
// (four groups = the four TMEM accumulator slots) turn 128 rows x 8 classes x 16 channels into bf16 CG8 stores, two
// parity classes (adjacent output voxels along d2) at a time so that a lane writes 32 contiguous bytes.
#include "conv_tc.cuh"
#include "tc_ptx.cuh"

#include <algorithm>
#include <stdlib.h>
#include <string.h>

namespace sgm {
namespace tc {

namespace {
using namespace tcptx;

constexpr int kIssuers = 2;
constexpr int kEpiGroups = 4;                                // == TMEM slots of 128 columns
constexpr int kThreads = (4 * kEpiGroups + 1 + kIssuers) * 32;
constexpr int kSmemMax = 227 * 1024;
constexpr int kPadPos = 64;
constexpr int kRingMax = 12;
constexpr int kN = 128;                                      // 8 parity classes x 16 output channels

struct PstArgs {
  int D[3];                 // INPUT extents; the output is 2 D
  int t1, t2, H1, H2, nt1, nt2;
  int H12, PS, m;           // positions of one plane slab (halo +1 on the high side), ring stride per group, tiles per slab
  uint32_t mH2;
  int nunits, units_per_win;
  int R;
  int cg0, cg1, cgA, act;   // channel groups of the two inputs; output channel groups of the TENSOR (window stride)
  int cg_first;             // first output channel group this launch writes (two groups = 16 channels per launch)
  float alpha;
  const __nv_bfloat16* w;
  const float* bias;
  __nv_bfloat16* out;
  int* error_flag;
  long long* trace;
};

template <int NCGP>
__global__ void __launch_bounds__(kThreads, 1)
pst_conv_kernel(const PstArgs a, const __grid_constant__ CUtensorMap tmap0, const __grid_constant__ CUtensorMap tmap1) {
  constexpr int NKB = 8 * NCGP;  // K blocks: input shift (s0, s1, s2) x 16 input channels
  constexpr int CG = 2 * NCGP;
  constexpr uint32_t W_BYTES = NKB * kN * 32;
  constexpr int PW = 4 * kEpiGroups;
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  uint8_t* w_smem = smem;
  uint8_t* a_region = smem + W_BYTES;
  uint8_t* ring = a_region + kPadPos * 16;
  const uint32_t plane_bytes = (uint32_t)CG * a.PS * 16u;
  const uint32_t a_bytes = (uint32_t)(kPadPos + 128 + kPadPos) * 16u + (uint32_t)a.R * plane_bytes;
  float* bias_s = reinterpret_cast<float*>(a_region + a_bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(bias_s + 16);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 1 + 2 * kRingMax + 8);
  const uint32_t bar0 = smem_u32(bars);
  const uint32_t WBAR = bar0;
  auto PFULL = [&](int s) { return bar0 + 8u * (1 + s); };
  auto PEMPTY = [&](int s) { return bar0 + 8u * (1 + kRingMax + s); };
  auto TFULL = [&](int s) { return bar0 + 8u * (1 + 2 * kRingMax + s); };
  auto TEMPTY = [&](int s) { return bar0 + 8u * (1 + 2 * kRingMax + 4 + s); };
  const int D0 = a.D[0], m = a.m, R = a.R;
  const int NP = D0 + 1;  // planes per unit: the plane behind the last one is outside the tensor (TMA zero fill)
  const bool tr = a.trace != nullptr && blockIdx.x == 0;

  if (tid == 0) {
    mbar_init(WBAR, 1);
    for (int s = 0; s < kRingMax; ++s) {
      mbar_init(PFULL(s), 1);
      mbar_init(PEMPTY(s), 2);  // the two tile rows (x0 = p - 1, x0 = p) that read plane p
    }
    for (int s = 0; s < 4; ++s) {
      mbar_init(TFULL(s), 1);
      mbar_init(TEMPTY(s), 4);  // the four warps of the epilogue group that owns the slot
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
  for (uint32_t i = tid; i < a_bytes / 16; i += kThreads) reinterpret_cast<uint4*>(a_region)[i] = make_uint4(0, 0, 0, 0);
  if (tid < 16) bias_s[tid] = __ldg(a.bias + tid);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int my_units = blockIdx.x < a.nunits ? (a.nunits - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

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
        for (int x0 = 0; x0 < NP; ++x0) {
          if (wrapped) mbar_wait_or_trap(PEMPTY(pslot), pphase ^ 1u, a.error_flag, 31);
          mbar_expect_tx(PFULL(pslot), (uint32_t)(CG * a.H12 * 16));
#pragma unroll
          for (int cg = 0; cg < CG; ++cg) {
            const bool first = cg < a.cg0;
            tma_load_4d(ring_base + (uint32_t)pslot * plane_bytes + (uint32_t)(cg * a.PS) * 16u, first ? &tmap0 : &tmap1,
                        (b2 * a.t2) * 8, b1 * a.t1, x0, first ? n * a.cg0 + cg : n * a.cg1 + (cg - a.cg0), PFULL(pslot));
          }
          if (++pslot == R) pslot = 0, pphase ^= 1u, wrapped = true;
        }
      }
    }
  } else if (warp > PW) {
    // ============================ MMA issuers: one elected lane each, alternating tile rows ============================
    if (elect_one()) {
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kN >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t ring16 = smem_u32(ring) >> 4;
      const uint32_t w_base16 = smem_u32(w_smem) >> 4;
      const uint32_t d_hi = 8u | (1u << 14);
      const uint32_t a_lbo = ((uint32_t)a.PS & 0x3FFFu) << 16;
      const uint32_t b_lbo = ((uint32_t)kN & 0x3FFFu) << 16;
      const uint32_t plane16 = (uint32_t)(CG * a.PS);
      mbar_wait_or_trap(WBAR, 0u, a.error_flag, 32);
      const int iw = warp - PW - 1;
      const int rows_total = my_units * D0;  // tile rows (one per input plane x0) this CTA sweeps
      // K-block operand offsets: block kb = (s0, s1, s2, cp); s0 selects the plane, (s1, s2) shift inside it
      uint32_t a_off[NKB], b_lo[NKB];
#pragma unroll
      for (int kb = 0; kb < NKB; ++kb) {
        const int s = kb / NCGP, cp = kb % NCGP;
        a_off[kb] = (uint32_t)(cp * 2 * a.PS + ((s >> 1) & 1) * a.H2 + (s & 1));
        b_lo[kb] = ((w_base16 + (uint32_t)(kb * kN * 2)) & 0x3FFFu) | b_lbo;
      }
      bool rdy_p = false, rdy_t = false;
      int stamps = 0;
      for (int rw = iw; rw < rows_total; rw += kIssuers) {
        const int u = rw / D0, x0 = rw - u * D0;
        const int pc = u * NP + x0;               // plane counter (producer order) of plane x0; plane x0 + 1 follows
        const int ps0 = pc % R, ps1 = (pc + 1) % R;
        const uint32_t ph1 = (uint32_t)((pc + 1) / R) & 1u;
        if (!rdy_p) mbar_wait_or_trap(PFULL(ps1), ph1, a.error_flag, 33);  // plane x0 landed before plane x0 + 1
        // ... but was it observed by THIS thread?  Planes arrive in order on one TMA queue and complete_tx of the
        // later plane implies the earlier one only per barrier: observe plane x0 as well (never blocks).
        mbar_wait_or_trap(PFULL(ps0), (uint32_t)(pc / R) & 1u, a.error_flag, 34);
        for (int j = 0; j < m; ++j) {
          const int T = rw * m + j;
          const int tslot = T & 3;
          if (!rdy_t && T >= 4) mbar_wait_or_trap(TEMPTY(tslot), ((uint32_t)(T >> 2) & 1u) ^ 1u, a.error_flag, 35);
          tc_fence_after();
          const uint32_t dcol = tmem_base + (uint32_t)(tslot * kN);
          const uint32_t base0 = (ring16 + (uint32_t)ps0 * plane16 + (uint32_t)(j * 128)) | a_lbo;
          const uint32_t base1 = (ring16 + (uint32_t)ps1 * plane16 + (uint32_t)(j * 128)) | a_lbo;
          constexpr int kSplit = NKB / 2;  // blocks [0, NKB/2) read plane x0 (s0 = 0), the rest plane x0 + 1
#pragma unroll
          for (int kb = 0; kb < kSplit; ++kb)
            tc_mma(dcol, ((uint64_t)d_hi << 32) | (base0 + a_off[kb]), ((uint64_t)d_hi << 32) | b_lo[kb], idesc, kb > 0 ? 1u : 0u);
          // poll the barriers of this issuer's next tile while the queued MMAs execute
          if (j + 1 < m) {
            const int Tn = T + 1;
            rdy_p = true;
            rdy_t = Tn < 4 || mbar_try_wait(TEMPTY(Tn & 3), ((uint32_t)(Tn >> 2) & 1u) ^ 1u);
          } else if (rw + kIssuers < rows_total) {
            const int rn = rw + kIssuers, un = rn / D0;
            const int pcn = un * NP + (rn - un * D0) + 1;
            const int Tn = rn * m;
            rdy_p = mbar_try_wait(PFULL(pcn % R), (uint32_t)(pcn / R) & 1u);
            rdy_t = Tn < 4 || mbar_try_wait(TEMPTY(Tn & 3), ((uint32_t)(Tn >> 2) & 1u) ^ 1u);
          }
#pragma unroll
          for (int kb = kSplit; kb < NKB; ++kb)
            tc_mma(dcol, ((uint64_t)d_hi << 32) | (base1 + a_off[kb]), ((uint64_t)d_hi << 32) | b_lo[kb], idesc, 1u);
          tc_commit(TFULL(tslot));
        }
        // plane p is read by tile rows p - 1 and p: every row releases both of its planes; the first / last plane of
        // a column has only one reader, which arrives a second time in place of the missing one
        tc_commit(PEMPTY(ps0));
        tc_commit(PEMPTY(ps1));
        if (x0 == 0) tc_commit(PEMPTY(ps0));
        if (x0 == D0 - 1) tc_commit(PEMPTY(ps1));
        if (tr && iw == 0 && stamps < 8) a.trace[1 + stamps++] = clock64();
      }
      if (tr && iw == 0) a.trace[12] = clock64();
    }
    __syncwarp();
  } else {
    // ============================ epilogue: group g owns TMEM slot g (tiles T with T % 4 == g) ============================
    const int egroup = warp >> 2, quarter = warp & 3;
    const uint32_t tlane = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(egroup * kN);
    const int OD1 = 2 * a.D[1], OD2 = 2 * a.D[2];
    const long long ovox = (long long)(2 * D0) * OD1 * OD2;
    const bool act = a.act != 0;
    const float alpha = a.alpha;
    float bias_r[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) bias_r[c] = bias_s[c];
    const int tiles_total = my_units * D0 * m;
    bool ok = true;
    for (int T = egroup; T < tiles_total && ok; T += kEpiGroups) {
      const int rw = T / m, j = T - rw * m;
      const int u = rw / D0, x0 = rw - u * D0;
      const int unit = blockIdx.x + u * gridDim.x;
      const int n = unit / a.units_per_win;
      const int r = unit - n * a.units_per_win;
      const int b1 = r / a.nt2, b2 = r - b1 * a.nt2;
      const int q = j * 128 + quarter * 32 + lane;
      const int h1 = (int)__umulhi((uint32_t)q, a.mH2);
      const int h2 = q - h1 * a.H2;
      const int r1 = b1 * a.t1 + h1, r2 = b2 * a.t2 + h2;
      const bool valid = q < a.H12 && h1 < a.t1 && h2 < a.t2 && r1 < a.D[1] && r2 < a.D[2];
      // output voxel of parity class (0, 0, 0); class (p0, p1, p2) adds (p0 * OD1 + p1) * OD2 + p2
      const long long obase = ((long long)(2 * x0) * OD1 + 2 * r1) * OD2 + 2 * r2;
      __nv_bfloat16* dst0 = a.out + (((long long)n * a.cgA + a.cg_first) * ovox + obase) * 8;
      ok = mbar_wait(TFULL(egroup), (uint32_t)(T >> 2) & 1u, a.error_flag, 36);
      if (!ok) break;
      tc_fence_after();
#pragma unroll
      for (int pp = 0; pp < 4; ++pp) {  // (p0, p1) pairs; the two p2 classes are adjacent columns blocks
        uint32_t raw[32];
        tc_ld16(tlane + (uint32_t)(pp * 32), raw);        // load + wait inside one asm statement
        tc_ld16(tlane + (uint32_t)(pp * 32 + 16), raw + 16);
        if (pp == 3) {  // the accumulator is in registers: hand the slot back to the issuers
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(TEMPTY(egroup));
        }
        if (!valid) continue;
        const int p0 = pp >> 1, p1 = pp & 1;
        __nv_bfloat16* dst = dst0 + ((long long)(p0 * OD1 + p1) * OD2) * 8;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          float v0[8], v1[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            float x = __uint_as_float(raw[8 * g + c]) + bias_r[8 * g + c];
            float y = __uint_as_float(raw[16 + 8 * g + c]) + bias_r[8 * g + c];
            if (act) x = prelu(x, alpha), y = prelu(y, alpha);
            v0[c] = x, v1[c] = y;
          }
          if (a.cg_first + g < a.cgA) {  // 32 contiguous bytes: output voxels 2 r2 and 2 r2 + 1
            uint4* d = reinterpret_cast<uint4*>(dst + (long long)g * ovox * 8);
            d[0] = pack8(v0);
            d[1] = pack8(v1);
          }
        }
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

struct PstPlan {
  int key[5];
  PstArgs args;
  int smem_bytes, grid;
};

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

int pst_fixed_smem(int ncgp) {
  return 8 * ncgp * kN * 32 + (kPadPos + 128 + kPadPos) * 16 + 16 * 4 + (1 + 2 * kRingMax + 8) * 8 + 16 + 128;
}

int pst_plan(const TcConv& c, const TcIO& io, PstPlan& pl) {
  PstArgs& a = pl.args;
  memset(&a, 0, sizeof(a));
  const int CG = 2 * c.pst_ncgp;
  for (int i = 0; i < 3; ++i) a.D[i] = io.id[i];
  const int nsm = sm_count();
  const int fixed = pst_fixed_smem(c.pst_ncgp);
  const double t_tile = 8.0 * c.pst_ncgp * (34.0 + 0.36 * kN);
  double best = 1e30;
  int bt1 = 0, bt2 = 0, bm = 0;
  for (int t1 = 1; t1 <= std::min(a.D[1], 62); ++t1)
    for (int t2 = 1; t2 <= std::min(a.D[2], 31); ++t2) {  // TMA box: H2 * 8 elements <= 256
      const int H1 = t1 + 1, H2 = t2 + 1;
      const int m = ceil_div(H1 * H2, 128);
      if (m > 2) continue;  // an accumulator slot must always be refilled by the same issuer (barrier phase tracking)
      const int PS = round_up(H1 * H2, 8);
      if (fixed + 4 * CG * PS * 16 > kSmemMax) continue;  // at least four planes in the ring
      const long long units = (long long)ceil_div(a.D[1], t1) * ceil_div(a.D[2], t2) * io.n;
      const double waves = (double)((units + nsm - 1) / nsm);
      const double cost = waves * ((double)a.D[0] * m * t_tile + 4000.0) * (1.0 + 0.02 / t2);
      if (cost < best) best = cost, bt1 = t1, bt2 = t2, bm = m;
    }
  SGM_REQUIRE(bt1 > 0, SGM_ERR_UNSUPPORTED, "pst_plan: no slab shape fits shared memory");
  a.t1 = bt1, a.t2 = bt2, a.H1 = bt1 + 1, a.H2 = bt2 + 1, a.m = bm;
  a.nt1 = ceil_div(a.D[1], bt1), a.nt2 = ceil_div(a.D[2], bt2);
  a.H12 = a.H1 * a.H2;
  a.PS = round_up(a.H12, 8);
  a.mH2 = (uint32_t)((0x100000000ULL + a.H2 - 1) / a.H2);
  a.units_per_win = a.nt1 * a.nt2;
  a.nunits = a.units_per_win * io.n;
  const int plane_bytes = CG * a.PS * 16;
  a.R = std::min(kRingMax, (kSmemMax - fixed) / plane_bytes);
  a.R = std::min(a.R, 8);
  pl.smem_bytes = fixed + a.R * plane_bytes;
  pl.grid = std::min(nsm, a.nunits);
  return SGM_OK;
}

template <int NCGP>
int launch_t(const PstArgs& a, const CUtensorMap& tm0, const CUtensorMap& tm1, int grid, int smem, cudaStream_t st) {
  static DeviceOnce attr_once;
  if (attr_once.first()) {
    SGM_CUDA_CHECK(cudaFuncSetAttribute(pst_conv_kernel<NCGP>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax));
  }
  pst_conv_kernel<NCGP><<<grid, kThreads, smem, st>>>(a, tm0, tm1);
  SGM_CUDA_CHECK(cudaGetLastError());
  return SGM_OK;
}

}  // namespace

int pst_pack(const sgm_conv_desc& d, TcConv* c) {
  c->pst_ncgp = 0;
  if (getenv("SGM_NO_PST")) return SGM_OK;
  if (d.kind != SGM_KIND_CONV_TRANSPOSE || d.kernel != 3 || d.stride != 2 || c->flat0 || c->mode != MODE_T2) return SGM_OK;
  if (d.cout > 32 || d.cin % 16 != 0) return SGM_OK;   // 17..32 output channels (11..32 classes): two launches of 16
  const int ncgp = d.cin / 16;
  if (ncgp != 2 && ncgp != 4) return SGM_OK;  // instantiated: 32 and 64 input channels
  const int NKB = 8 * ncgp;
  const int npass = (d.cout + 15) / 16;
  const size_t pass_elems = (size_t)NKB * 2 * kN * 8;
  std::vector<uint16_t> w(pass_elems * npass, 0);
  for (int pass = 0; pass < npass; ++pass)
  for (int s = 0; s < 8; ++s)
    for (int cp = 0; cp < ncgp; ++cp) {
      const int kb = s * ncgp + cp;
      const int sh[3] = {(s >> 2) & 1, (s >> 1) & 1, s & 1};
      for (int cls = 0; cls < 8; ++cls) {
        int kk[3];
        bool used = true;
        for (int ax = 0; ax < 3; ++ax) {
          const int pbit = (cls >> (2 - ax)) & 1;
          if (pbit == 0) kk[ax] = 1, used = used && sh[ax] == 0;  // even outputs read shift 0 only (centre tap)
          else kk[ax] = sh[ax] == 0 ? 2 : 0;
        }
        if (!used) continue;
        const int tap = (kk[0] * 3 + kk[1]) * 3 + kk[2];
        for (int kc = 0; kc < 2; ++kc)
          for (int cl = 0; cl < 16 && pass * 16 + cl < d.cout; ++cl)
            for (int k8 = 0; k8 < 8; ++k8) {
              const int ci = (cp * 2 + kc) * 8 + k8, co = pass * 16 + cl;
              // ConvTranspose weight layout [Cin][Cout][27]
              w[pass * pass_elems + (((size_t)kb * 2 + kc) * kN + (cls * 16 + cl)) * 8 + k8] =
                  f2bf(d.weight[((size_t)ci * d.cout + co) * 27 + tap]);
            }
      }
    }
  if (cudaMalloc(&c->pst_w, w.size() * 2) != cudaSuccess) {
    set_error("pst_pack: cudaMalloc failed: %s", cudaGetErrorString(cudaGetLastError()));
    return SGM_ERR_CUDA;
  }
  SGM_CUDA_CHECK(cudaMemcpy(c->pst_w, w.data(), w.size() * 2, cudaMemcpyHostToDevice));
  c->pst_ncgp = ncgp;
  c->pst_npass = npass;
  c->pst_plan_cache = new std::vector<PstPlan>();
  return SGM_OK;
}

void pst_free(TcConv* c) {
  if (c->pst_w) cudaFree(c->pst_w);
  c->pst_w = nullptr;
  delete reinterpret_cast<std::vector<PstPlan>*>(c->pst_plan_cache);
  c->pst_plan_cache = nullptr;
}

bool pst_applicable(const TcConv& c, const TcIO& io) {
  if (!c.pst_ncgp || !c.pst_w || !tma_available()) return false;
  if (io.outB || io.res || io.out_kind != OUT_CG8) return false;
  if (io.cg0 + io.cg1 != 2 * c.pst_ncgp || io.cg0 < 1) return false;
  if (io.cg1 > 0 && !io.in1) return false;
  for (int i = 0; i < 3; ++i)
    if (io.od[i] != 2 * io.id[i]) return false;
  if ((long long)io.od[0] * io.od[1] * io.od[2] >= (1LL << 31)) return false;
  return true;
}

int pst_launch(const TcConv& c, const TcIO& io, int* error_flag_dev, cudaStream_t st) {
  auto* plans = reinterpret_cast<std::vector<PstPlan>*>(c.pst_plan_cache);
  const int key[5] = {io.id[0], io.id[1], io.id[2], io.n, 0};
  const PstPlan* pe = nullptr;
  for (auto& e : *plans)
    if (memcmp(e.key, key, sizeof(key)) == 0) pe = &e;
  if (!pe) {
    PstPlan e;
    memcpy(e.key, key, sizeof(key));
    int rc = pst_plan(c, io, e);
    if (rc) return rc;
    plans->push_back(e);
    pe = &plans->back();
  }
  PstArgs a = pe->args;
  a.cg0 = io.cg0, a.cg1 = io.cg1, a.cgA = io.cgA, a.act = c.actA, a.alpha = c.alphaA;
  a.w = c.pst_w, a.bias = c.bias;
  a.out = (__nv_bfloat16*)io.outA;
  a.error_flag = error_flag_dev;
  static const bool dbg = getenv("SGM_DEBUG") != nullptr;
  if (dbg)
    fprintf(stderr, "[pst_launch] NCGP=%d D=(%d,%d,%d) n=%d t=(%d,%d) H12=%d PS=%d m=%d units=%d grid=%d smem=%d R=%d\n",
            c.pst_ncgp, a.D[0], a.D[1], a.D[2], io.n, a.t1, a.t2, a.H12, a.PS, a.m, a.nunits, pe->grid, pe->smem_bytes, a.R);
  static const bool trace_on = getenv("SGM_TRACE") != nullptr;
  static long long* trace_dev = nullptr;
  if (trace_on) {
    if (!trace_dev) cudaMalloc(&trace_dev, 16 * sizeof(long long));
    cudaMemsetAsync(trace_dev, 0, 16 * sizeof(long long), st);
    a.trace = trace_dev;
  }
  CUtensorMap tm0, tm1;
  const int box[3] = {1, a.H1, a.H2};
  int rc = make_brick_map(&tm0, io.in0, io.n * io.cg0, io.id, box);
  if (rc) return rc;
  tm1 = tm0;
  if (io.cg1 > 0) {
    rc = make_brick_map(&tm1, io.in1, io.n * io.cg1, io.id, box);
    if (rc) return rc;
  }
  for (int pass = 0; pass < c.pst_npass; ++pass) {  // 16 output channels per launch
    a.cg_first = 2 * pass;
    a.w = c.pst_w + (size_t)pass * 8 * c.pst_ncgp * 2 * kN * 8;
    a.bias = c.bias + 16 * pass;
    if (c.pst_ncgp == 2) rc = launch_t<2>(a, tm0, tm1, pe->grid, pe->smem_bytes, st);
    else rc = launch_t<4>(a, tm0, tm1, pe->grid, pe->smem_bytes, st);
    if (rc) return rc;
  }
  if (trace_on) {
    long long t[16];
    cudaStreamSynchronize(st);
    cudaMemcpy(t, trace_dev, sizeof(t), cudaMemcpyDeviceToHost);
    auto d = [&](int i) { return t[i] ? (double)(t[i] - t[0]) : -1.0; };
    fprintf(stderr,
            "[pst trace] NCGP=%d D=(%d,%d,%d) n=%d t=(%d,%d) m=%d units/cta=%.1f | rows issued %.0f %.0f %.0f %.0f %.0f %.0f "
            "%.0f %.0f | issuer done %.0f, end %.0f cycles\n",
            c.pst_ncgp, a.D[0], a.D[1], a.D[2], io.n, a.t1, a.t2, a.m, (double)a.nunits / pe->grid, d(1), d(2), d(3), d(4), d(5),
            d(6), d(7), d(8), d(12), d(9));
  }
  return SGM_OK;
}

}  // namespace tc
}  // namespace sgm
