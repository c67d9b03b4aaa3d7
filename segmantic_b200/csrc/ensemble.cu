// Ensemble combination of several models' predictions (seg/monai_unet.py:848-1004: MeanEnsembled + AsDiscreted,
// AsDiscreted + VoteEnsembled, AsDiscreted + SelectBestEnsembled from seg/transforms.py:15-61).  Voxel-wise, HBM-bound
// streaming kernels over stacked model outputs.
#include "common.cuh"

#include <algorithm>

namespace sgm {
namespace {

constexpr int kMaxModels = 16;
constexpr int kMaxEnsClasses = 64;

struct EnsArgs {
  int n_models, num_classes;
  long long voxels;
  float scale[kMaxModels];  // mean: w_m / mean(w)
  int tissue[kMaxEnsClasses], model[kMaxEnsClasses];  // select_best: (tissue id -> model index), in dictionary order
  int n_pairs;
};

// MONAI MeanEnsemble(weights) + AsDiscrete(argmax): out[c] = mean_m(x_m[c] * w_m / mean(w)), label = argmax_c (ties ->
// lowest class).  logits: [E][C][V] float32.
__global__ void __launch_bounds__(256) ens_mean_argmax_kernel(const float* __restrict__ logits, uint8_t* __restrict__ labels,
                                                              float* __restrict__ mean_out, const EnsArgs a) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long cv = (long long)a.num_classes * a.voxels;
  const float inv_e = 1.f / (float)a.n_models;
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < a.voxels; v += stride) {
    float best = 0.f;
    int arg = 0;
    for (int c = 0; c < a.num_classes; ++c) {
      float s = 0.f;
      for (int m = 0; m < a.n_models; ++m) s = __fadd_rn(s, __fmul_rn(__ldcs(logits + m * cv + c * a.voxels + v), a.scale[m]));
      s = __fmul_rn(s, inv_e);
      if (mean_out) __stcs(mean_out + c * a.voxels + v, s);
      if (c == 0 || s > best) best = s, arg = c;
    }
    labels[v] = (uint8_t)arg;
  }
}

// MONAI VoteEnsemble(num_classes) on discrete single-channel predictions: one-hot, mean over models, argmax (ties ->
// lowest class).  labels_in: [E][V] uint8.
__global__ void __launch_bounds__(256) ens_vote_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, const EnsArgs a) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < a.voxels; v += stride) {
    uint8_t l[kMaxModels];
    for (int m = 0; m < a.n_models; ++m) l[m] = in[m * a.voxels + v];
    int best_cnt = 0, best = 0;  // no vote inside [0, num_classes) at all -> class 0 (all-zero one-hot rows)
    for (int m = 0; m < a.n_models; ++m) {
      const int cls = l[m];
      if (cls >= a.num_classes) continue;
      int cnt = 0;
      for (int k = 0; k < a.n_models; ++k) cnt += l[k] == cls;
      if (cnt > best_cnt || (cnt == best_cnt && cls < best)) best_cnt = cnt, best = cls;
    }
    out[v] = (uint8_t)best;
  }
}

// SelectBestEnsemble (seg/transforms.py:40-52): for (tissue, model) in dictionary order, voxels the chosen model labels
// `tissue` get `tissue` (later pairs overwrite earlier ones); voxels no pair claims are 0 (the reference leaves them
// uninitialised -- torch.empty).
__global__ void __launch_bounds__(256) ens_select_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, const EnsArgs a) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < a.voxels; v += stride) {
    int r = 0;
    for (int p = 0; p < a.n_pairs; ++p)
      if (in[a.model[p] * a.voxels + v] == a.tissue[p]) r = a.tissue[p];
    out[v] = (uint8_t)r;
  }
}

int ens_grid(long long n) { return (int)std::max<long long>(1, std::min<long long>((n + 255) / 256, 148LL * 16)); }

}  // namespace
}  // namespace sgm

using namespace sgm;

extern "C" int32_t sgm_ensemble_mean_argmax(const float* logits_dev, int32_t n_models, int32_t num_classes, int64_t voxels,
                                            const float* weights, uint8_t* labels_dev, float* mean_dev, void* stream) {
  SGM_REQUIRE(logits_dev && labels_dev && voxels >= 0, SGM_ERR_INVALID, "ensemble_mean_argmax: bad argument");
  SGM_REQUIRE(n_models >= 1 && n_models <= kMaxModels && num_classes >= 1 && num_classes <= 255, SGM_ERR_UNSUPPORTED,
              "ensemble_mean_argmax: 1..%d models and 1..255 classes, got %d / %d", kMaxModels, n_models, num_classes);
  EnsArgs a = {};
  a.n_models = n_models, a.num_classes = num_classes, a.voxels = voxels;
  float wbar = 0.f;
  for (int m = 0; m < n_models; ++m) wbar += weights ? weights[m] : 1.f;
  wbar /= (float)n_models;
  SGM_REQUIRE(wbar != 0.f, SGM_ERR_INVALID, "ensemble_mean_argmax: weights sum to zero");
  for (int m = 0; m < n_models; ++m) a.scale[m] = (weights ? weights[m] : 1.f) / wbar;
  if (voxels == 0) return SGM_OK;
  ens_mean_argmax_kernel<<<ens_grid(voxels), 256, 0, (cudaStream_t)stream>>>(logits_dev, labels_dev, mean_dev, a);
  SGM_CUDA_CHECK(cudaGetLastError());
  return SGM_OK;
}

extern "C" int32_t sgm_ensemble_vote(const uint8_t* labels_in_dev, int32_t n_models, int32_t num_classes, int64_t voxels,
                                     uint8_t* labels_out_dev, void* stream) {
  SGM_REQUIRE(labels_in_dev && labels_out_dev && voxels >= 0, SGM_ERR_INVALID, "ensemble_vote: bad argument");
  SGM_REQUIRE(n_models >= 1 && n_models <= kMaxModels && num_classes >= 1 && num_classes <= 256, SGM_ERR_UNSUPPORTED,
              "ensemble_vote: 1..%d models, got %d", kMaxModels, n_models);
  EnsArgs a = {};
  a.n_models = n_models, a.num_classes = num_classes, a.voxels = voxels;
  if (voxels == 0) return SGM_OK;
  ens_vote_kernel<<<ens_grid(voxels), 256, 0, (cudaStream_t)stream>>>(labels_in_dev, labels_out_dev, a);
  SGM_CUDA_CHECK(cudaGetLastError());
  return SGM_OK;
}

extern "C" int32_t sgm_ensemble_select_best(const uint8_t* labels_in_dev, int32_t n_models, int64_t voxels,
                                            const int32_t* tissue_ids, const int32_t* model_ids, int32_t n_pairs,
                                            uint8_t* labels_out_dev, void* stream) {
  SGM_REQUIRE(labels_in_dev && labels_out_dev && voxels >= 0 && tissue_ids && model_ids, SGM_ERR_INVALID,
              "ensemble_select_best: bad argument");
  SGM_REQUIRE(n_models >= 1 && n_models <= kMaxModels && n_pairs >= 0 && n_pairs <= kMaxEnsClasses, SGM_ERR_UNSUPPORTED,
              "ensemble_select_best: 1..%d models and <= %d (tissue, model) pairs", kMaxModels, kMaxEnsClasses);
  EnsArgs a = {};
  a.n_models = n_models, a.voxels = voxels, a.n_pairs = n_pairs;
  for (int p = 0; p < n_pairs; ++p) {
    SGM_REQUIRE(model_ids[p] >= 0 && model_ids[p] < n_models && tissue_ids[p] >= 0 && tissue_ids[p] <= 255, SGM_ERR_INVALID,
                "ensemble_select_best: pair %d = (tissue %d, model %d) out of range", p, tissue_ids[p], model_ids[p]);
    a.tissue[p] = tissue_ids[p], a.model[p] = model_ids[p];
  }
  if (voxels == 0) return SGM_OK;
  ens_select_kernel<<<ens_grid(voxels), 256, 0, (cudaStream_t)stream>>>(labels_in_dev, labels_out_dev, a);
  SGM_CUDA_CHECK(cudaGetLastError());
  return SGM_OK;
}
