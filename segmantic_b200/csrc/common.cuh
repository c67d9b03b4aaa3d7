// Shared declarations of the segmantic_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/segmantic_b200.h"

namespace sgm {

void set_error(const char* fmt, ...);

#define SGM_CUDA_CHECK(expr)                                                              \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      sgm::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,    \
                     __LINE__);                                                           \
      return SGM_ERR_CUDA;                                                                \
    }                                                                                     \
  } while (0)

#define SGM_REQUIRE(cond, code, ...)  \
  do {                                \
    if (!(cond)) {                    \
      sgm::set_error(__VA_ARGS__);    \
      return (code);                  \
    }                                 \
  } while (0)

// Activation tensors between layers use the "CG8" layout: [n][cg][d0][d1][d2][8 channels], i.e.
// 8-channel groups interleaved per voxel (32 B per voxel-group in fp32, 16 B in bf16).  A voxel's
// group is one vector load; consecutive d2 voxels are contiguous, so position-per-lane kernels
// coalesce, and in bf16 the 16-byte groups are exactly the K-major "core matrix" rows tcgen05 wants.
struct ConvArgs {
  // input: concatenation of up to two CG8 tensors along channels, or (IN_PLANAR) a planar fp32
  // volume [cin][vd0][vd1][vd2] from which window n is read at origin win_origin[n].
  const void* in0;
  const void* in1;
  int cg0, cg1;
  int cin_real;
  int n;
  int id[3], od[3];
  int k[3], s[3], pad[3];
  const void* w;     // packed weights (layout depends on the kernel family)
  const float* bias; // [cout_groups*8], zero padded
  int cout_groups;
  void* out;         // CG8 [n][cout_groups][od]
  const void* res;   // CG8 residual added after the activation, or nullptr
  int act;
  float alpha;
  // planar input (stem)
  long long vol_cstride;
  int vd1, vd2;
  const int* win_origin; // device int[n][3], in vol-buffer coordinates
  // planar output (head): OUT_PLANAR stores logits [n][c_real][od]; OUT_BLEND accumulates
  // acc[c][x][y][z] = acc + logit * imap for window `wo`, planes clipped to [0, ad0).
  float* pl_out;
  long long pl_cstride; // channel stride of pl_out (= voxels of the target volume / roi)
  long long pl_nstride; // batch stride (OUT_PLANAR)
  int ad0, ad1, ad2;
  int wo[3];
  const float* imap0;
  const float* imap1;
  const float* imap2;
  float imap_floor;
  int c_real;
  int pl_weighted;  // OUT_PLANAR: logits pre-multiplied by the importance map (deferred blend)
};

enum { OUT_CG8 = 0, OUT_BLEND = 1, OUT_PLANAR = 2 };

// ---- fp32 CUDA-core family (conv_fp32.cu)
int fp32_conv_cout_tile(bool transposed);
int launch_conv_fp32(const ConvArgs& a, bool bf16_storage, bool in_planar, int out_kind, cudaStream_t st);
int launch_convT_fp32(const ConvArgs& a, bool bf16_storage, cudaStream_t st);
// fused stem (first down block's unit0 + residual-branch conv, planar fp32 input)
int stem_cout_tile();
int launch_stem(const ConvArgs& a, int cgA, int cgB, void* outA, void* outB, bool bf16_storage, cudaStream_t st);

// tensor-core stem (conv_stem_tc.cu): 3-D stride-2 k3 first block with 1-2 input channels, bf16 precision
int stem_tc_pack(const sgm_conv_desc& u0, const sgm_conv_desc& rs, int spatial_dims, void** w_dev, float** b_dev, int* kp);
int launch_stem_tc(const ConvArgs& a, const void* w_dev, const float* b_dev, int kp, int cgA, int cgB, void* outA,
                   void* outB, int* error_flag_dev, cudaStream_t st);

// ---- bf16 tcgen05 family (conv_tc.cu)
struct TcConvPlan;

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Process-wide lock for the library's few caches (tensor maps, K-block tables): handles are per device, but one
// process may drive several devices from several threads (predict(gpu_ids=[0, 1, ...])).
void lock_global();
void unlock_global();
struct GlobalLock {
  GlobalLock() { lock_global(); }
  ~GlobalLock() { unlock_global(); }
};

// One-time set-up per DEVICE at a call site (cudaFuncSetAttribute and __constant__ uploads apply to the current device
// only):  static DeviceOnce once;  if (once.first()) { ... }
struct DeviceOnce {
  unsigned long long mask = 0;
  bool first() {
    int dev = 0;
    cudaGetDevice(&dev);
    GlobalLock g;
    const unsigned long long bit = 1ull << (dev & 63);
    if (mask & bit) return false;
    mask |= bit;
    return true;
  }
};

}  // namespace sgm
