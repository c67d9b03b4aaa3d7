// Evaluation branch of predict() (seg/monai_unet.py:640-725, seg/evaluation.py:96-125): the confusion matrix of a
// predicted label map against a ground-truth label map -- a histogram over (y, y_pred) pairs from which Dice and
// MONAI's confusion-matrix metrics follow on the host.  HBM-bound: 2 bytes per voxel read once; per-CTA counters in
// shared memory (num_classes^2 <= 4096 cells), flushed with one global atomic per non-empty cell.
#include "common.cuh"

#include <algorithm>

namespace sgm {
namespace {

constexpr int kMaxClasses = 64;

__global__ void __launch_bounds__(256) confusion_kernel(const uint8_t* __restrict__ y_pred, const uint8_t* __restrict__ y,
                                                        long long n, int num_classes, unsigned long long* cm,
                                                        unsigned long long* ignored) {
  extern __shared__ unsigned int cells[];  // [num_classes][num_classes], row = true label, column = prediction
  __shared__ unsigned int skipped;
  const int ncell = num_classes * num_classes;
  for (int i = threadIdx.x; i < ncell; i += blockDim.x) cells[i] = 0u;
  if (threadIdx.x == 0) skipped = 0u;
  __syncthreads();
  const long long nvec = n / 16;  // 16 label pairs per thread and iteration (two 16-byte loads)
  const long long stride = (long long)gridDim.x * blockDim.x;
  const bool aligned = ((reinterpret_cast<uintptr_t>(y_pred) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
  unsigned int bad = 0;
  auto count = [&](unsigned int t, unsigned int p) {
    if (t < (unsigned)num_classes && p < (unsigned)num_classes) atomicAdd(&cells[t * num_classes + p], 1u);
    else ++bad;
  };
  long long done = 0;
  if (aligned) {
    for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
      const uint4 a = __ldcs(reinterpret_cast<const uint4*>(y) + v);
      const uint4 b = __ldcs(reinterpret_cast<const uint4*>(y_pred) + v);
      const unsigned int aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int w = 0; w < 4; ++w)
#pragma unroll
        for (int k = 0; k < 4; ++k) count((aw[w] >> (8 * k)) & 255u, (bw[w] >> (8 * k)) & 255u);
    }
    done = nvec * 16;
  }
  for (long long i = done + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) count(y[i], y_pred[i]);
  if (bad) atomicAdd(&skipped, bad);
  __syncthreads();
  for (int i = threadIdx.x; i < ncell; i += blockDim.x)
    if (cells[i]) atomicAdd(cm + i, (unsigned long long)cells[i]);
  if (threadIdx.x == 0 && skipped && ignored) atomicAdd(ignored, (unsigned long long)skipped);
}

}  // namespace
}  // namespace sgm

using namespace sgm;

extern "C" int32_t sgm_confusion_matrix(const uint8_t* y_pred_dev, const uint8_t* y_dev, int64_t voxels,
                                        int32_t num_classes, int64_t* cm_dev, int64_t* ignored_dev, void* stream) {
  SGM_REQUIRE(cm_dev && voxels >= 0 && (voxels == 0 || (y_pred_dev && y_dev)), SGM_ERR_INVALID, "confusion_matrix: bad argument");
  SGM_REQUIRE(num_classes >= 1 && num_classes <= kMaxClasses, SGM_ERR_UNSUPPORTED,
              "confusion_matrix supports 1..%d classes, got %d", kMaxClasses, num_classes);
  cudaStream_t st = (cudaStream_t)stream;
  SGM_CUDA_CHECK(cudaMemsetAsync(cm_dev, 0, sizeof(int64_t) * num_classes * num_classes, st));
  if (ignored_dev) SGM_CUDA_CHECK(cudaMemsetAsync(ignored_dev, 0, sizeof(int64_t), st));
  if (voxels == 0) return SGM_OK;
  const long long work = (voxels + 15) / 16;
  const int blocks = (int)std::max<long long>(1, std::min<long long>((work + 255) / 256, 148LL * 8));
  confusion_kernel<<<blocks, 256, sizeof(unsigned int) * num_classes * num_classes, st>>>(
      y_pred_dev, y_dev, voxels, num_classes, reinterpret_cast<unsigned long long*>(cm_dev),
      reinterpret_cast<unsigned long long*>(ignored_dev));
  SGM_CUDA_CHECK(cudaGetLastError());
  return SGM_OK;
}
