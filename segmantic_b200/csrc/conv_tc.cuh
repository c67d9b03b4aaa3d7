// tcgen05 implicit-GEMM convolution family (bf16 operands, fp32 accumulation in TMEM), sm_100a.
#pragma once
#include "common.cuh"

#include <cuda.h>  // CUtensorMap (types only; the encoder is resolved through cudaGetDriverEntryPoint)

#include <vector>

namespace sgm {
namespace tc {

enum { MODE_S1 = 0, MODE_S2 = 1, MODE_T2 = 2 };

// One K=16 step of the implicit GEMM: which 16-byte-row view of the shared-memory activation
// slab is the A operand, and which accumulator (output parity class) it feeds.
struct KBlock {
  int slab;      // parity slab (S2) or 0
  int cgpair;    // pair of input channel groups (16 channels)
  int shift[3];  // row-space shift per axis
  int cls;       // accumulator class (T2) or 0
  int first;     // first block of its class -> overwrite instead of accumulate
};

// Device-resident, per-convolution packing for the tcgen05 family.
struct TcConv {
  int mode = MODE_S1;
  int cin = 0, cgin = 0;      // cgin even (channels padded to 16)
  int ntot = 0;               // fused output channels, padded to 16 (main [+ residual branch])
  int ncta = 0;               // output channels per CTA (MMA N)
  int ncoblk = 0;
  int k[3] = {3, 3, 3};       // kernel extent per axis (1 on the flat axis of 2-D networks)
  int flat0 = 0;              // 2-D network: axis 0 has extent 1, kernel 1, stride 1
  int ncls = 1;
  int kmajor = 0;             // deep stride-1 layer: channel-major K blocks, brick streamed in channel chunks
  int tfold = 0;              // transposed conv: parity classes folded into the MMA N dimension
  int mma_n = 0;              // MMA N (ncta, or ncls * ncta when tfold)
  std::vector<KBlock> blocks; // host copy, in weight-pack order
  __nv_bfloat16* w = nullptr; // device [ncoblk][nblk][2][ncta][8]
  float* bias = nullptr;      // device [ntot]
  // epilogue segments in units of output channel groups of the fused N
  int segA_cg = 0;            // first segA_cg groups -> out A (act_a / alpha_a), rest -> out B (no act)
  int actA = 0;
  float alphaA = 0.f;
  int c_real = 0;             // real (unpadded) channels of segment A
  int blk_off = -1;           // slice of the constant K-block bank
  void* plan_cache = nullptr; // launch geometry per (dims, batch), filled lazily by tc_launch
  // "plane-sweep" packing (conv_ps.cu): stride-1 3x3x3 convs with <= 32 channels fold the three d0 taps
  // into the MMA N dimension (9 instead of 27 MMAs per 16 input channels) and sweep the window along d0
  __nv_bfloat16* ps_w = nullptr;  // device [9*ncgp][2][NP][8], row n' = k0*CB + co
  int ps_cb = 0;                  // columns per d0 tap (4, 8, 10, 16, 20, 24 or 32); 0 = not eligible
  int ps_ncgp = 0;                // input channel pairs of groups (cin padded to 16*ncgp)
  void* ps_plan_cache = nullptr;
  // transposed plane-sweep packing (conv_pst.cu): stride-2 transposed convs with <= 16 output channels fold the 8
  // output parity classes into the MMA N dimension
  __nv_bfloat16* pst_w = nullptr;  // device [8*ncgp][2][128][8], row n' = class*16 + co
  int pst_ncgp = 0;                // input channel pairs of groups (2 or 4); 0 = not eligible
  int pst_npass = 1;               // launches of 16 output channels each (2 for 17..32 output channels)
  void* pst_plan_cache = nullptr;
  // row-sweep packing (conv_rs.cu): the head (<= 16 -> <= 10 channels, planar fp32 output) folds the d0 AND d1 taps
  // into the MMA N dimension through overlapping accumulator columns
  __nv_bfloat16* rs_w = nullptr;   // device [3 rotations][3 k2][2][9*CP+16][8]
  int rs_cp = 0;                   // accumulator columns per (output row, plane slot); 0 = not eligible
  void* rs_plan_cache = nullptr;
  // channel-streamed persistent packing (conv_cs.cu): K-heavy 3-D layers (>= 64 input channels; strided from 32)
  __nv_bfloat16* cs_w = nullptr;   // device [coblk][16-channel chunk][tap block][2][NB][8]
  void* cs_state = nullptr;        // CsState (packing geometry + launch plans); null = not eligible
};

struct TcIO {
  const void* in0 = nullptr;  // bf16 CG8
  const void* in1 = nullptr;
  int cg0 = 0, cg1 = 0;
  int n = 1;
  int id[3] = {1, 1, 1};
  int od[3] = {1, 1, 1};
  void* outA = nullptr;       // bf16 CG8 [n][cgA][od]
  int cgA = 0;
  void* outB = nullptr;       // bf16 CG8 [n][cgB][od] (fused residual branch) or nullptr
  int cgB = 0;
  const void* res = nullptr;  // bf16 CG8 residual added to segment A after the activation
  int out_kind = OUT_CG8;     // OUT_CG8 | OUT_BLEND | OUT_PLANAR (segment A only)
  float* pl_out = nullptr;
  int pl_weighted = 0;        // OUT_PLANAR: logits pre-multiplied by the importance map (deferred blend)
  long long pl_cstride = 0, pl_nstride = 0;
  int ad0 = 0, ad1 = 0, ad2 = 0;
  int wo[3] = {0, 0, 0};
  const float* imap[3] = {nullptr, nullptr, nullptr};
  float imap_floor = 0.f;
};

// Build the packing for a convolution (optionally fused with a second conv reading the same input).
// `main` and `second` are folded fp32 descriptors; `second` may be null.
int tc_pack(const sgm_conv_desc* main_desc, const sgm_conv_desc* second, int spatial_dims, TcConv** out);
void tc_free(TcConv* c);
bool tc_supported(const sgm_conv_desc& d);
int tc_launch(const TcConv& c, const TcIO& io, int* error_flag_dev, cudaStream_t st);

// shared host helpers (conv_tc.cu)
bool tma_available();
int make_brick_map(CUtensorMap* out, const void* ptr, int ncg, const int d[3], const int H[3],
                   const int* par = nullptr);

// plane-sweep family (conv_ps.cu)
int ps_pack(const sgm_conv_desc& d, TcConv* c);          // fills ps_* when the conv is eligible (else leaves ps_cb = 0)
bool ps_applicable(const TcConv& c, const TcIO& io);
int ps_launch(const TcConv& c, const TcIO& io, int* error_flag_dev, cudaStream_t st);
void ps_free(TcConv* c);

// row-sweep family (conv_rs.cu)
int rs_pack(const sgm_conv_desc& d, TcConv* c);
bool rs_applicable(const TcConv& c, const TcIO& io);
int rs_launch(const TcConv& c, const TcIO& io, int* error_flag_dev, cudaStream_t st);
void rs_free(TcConv* c);

// channel-streamed persistent family (conv_cs.cu)
int cs_pack(const sgm_conv_desc* main_desc, const sgm_conv_desc* second, TcConv* c);
bool cs_applicable(const TcConv& c, const TcIO& io);
int cs_launch(const TcConv& c, const TcIO& io, int* error_flag_dev, cudaStream_t st);
void cs_free(TcConv* c);

// transposed plane-sweep family (conv_pst.cu)
int pst_pack(const sgm_conv_desc& d, TcConv* c);
bool pst_applicable(const TcConv& c, const TcIO& io);
int pst_launch(const TcConv& c, const TcIO& io, int* error_flag_dev, cudaStream_t st);
void pst_free(TcConv* c);

}  // namespace tc
}  // namespace sgm
