// NVLink peer-memory exchange for the multi-GPU driver (one process per GPU; SURVEY.md section 8e, option 2).
//
// The window-ownership partition needs ONE data-path exchange per seam: the importance-weighted logits of the last
// windows of rank r also cover the first output planes of rank r + 1.  Instead of an NCCL send / recv pair (whose
// kernels compete for SMs with the persistent conv kernels and serialise on the communicator), rank r + 1 exports its
// receive buffer as a CUDA IPC handle, rank r maps it and PUSHES the windows with the copy engines
// (cudaMemcpyAsync device-to-device over NVLink: no SM involved), then raises a flag in the peer's memory; rank r + 1
// waits for the flag on its own stream before it blends, and acknowledges in rank r's memory when the blend has read
// the buffer, so that the next volume's push cannot overtake it.  torch.distributed is used for the handle exchange
// and the final label gather only.
#include "common.cuh"

#include <cuda.h>
#include <string.h>

namespace sgm {
namespace {

typedef CUresult (*GetAddressRangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);

GetAddressRangeFn address_range_fn() {
  static GetAddressRangeFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (GetAddressRangeFn)p;
  }
  return fn;
}

__global__ void p2p_signal_kernel(uint32_t* flag, uint32_t value) {
  // everything ordered before this kernel on the stream (the pushed data) is complete; publish system-wide
  __threadfence_system();
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(value) : "memory");
}

__global__ void p2p_wait_kernel(const uint32_t* flag, uint32_t value, int32_t* timeout_flag, long long max_cycles) {
  const long long t0 = clock64();
  for (;;) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
    if ((int32_t)(v - value) >= 0) break;  // counters only grow (wrap-around safe comparison)
    if (clock64() - t0 > max_cycles) {
      if (timeout_flag) atomicExch(timeout_flag, 1);
      break;
    }
    __nanosleep(200);
  }
}

}  // namespace
}  // namespace sgm

using namespace sgm;

extern "C" int32_t sgm_p2p_export(const void* ptr_dev, uint8_t handle[64], int64_t* offset) {
  SGM_REQUIRE(ptr_dev && handle && offset, SGM_ERR_INVALID, "sgm_p2p_export: null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  GetAddressRangeFn range = address_range_fn();
  SGM_REQUIRE(range, SGM_ERR_CUDA, "sgm_p2p_export: cuMemGetAddressRange unavailable");
  CUdeviceptr base = 0;
  size_t size = 0;
  const CUresult r = range(&base, &size, (CUdeviceptr)ptr_dev);
  SGM_REQUIRE(r == CUDA_SUCCESS, SGM_ERR_CUDA, "sgm_p2p_export: cuMemGetAddressRange failed (%d)", (int)r);
  cudaIpcMemHandle_t h;
  SGM_CUDA_CHECK(cudaIpcGetMemHandle(&h, (void*)base));
  memcpy(handle, &h, 64);
  *offset = (int64_t)((CUdeviceptr)ptr_dev - base);
  return SGM_OK;
}

extern "C" int32_t sgm_p2p_open(const uint8_t handle[64], void** base_out) {
  SGM_REQUIRE(handle && base_out, SGM_ERR_INVALID, "sgm_p2p_open: null argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  *base_out = nullptr;
  SGM_CUDA_CHECK(cudaIpcOpenMemHandle(base_out, h, cudaIpcMemLazyEnablePeerAccess));
  return SGM_OK;
}

extern "C" int32_t sgm_p2p_close(void* base) {
  if (!base) return SGM_OK;
  SGM_CUDA_CHECK(cudaIpcCloseMemHandle(base));
  return SGM_OK;
}

extern "C" int32_t sgm_p2p_put(void* dst_peer_dev, const void* src_dev, int64_t bytes, void* stream) {
  SGM_REQUIRE(dst_peer_dev && src_dev && bytes >= 0, SGM_ERR_INVALID, "sgm_p2p_put: bad argument");
  if (bytes == 0) return SGM_OK;
  SGM_CUDA_CHECK(cudaMemcpyAsync(dst_peer_dev, src_dev, (size_t)bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return SGM_OK;
}

extern "C" int32_t sgm_p2p_signal(uint32_t* flag_peer_dev, uint32_t value, void* stream) {
  SGM_REQUIRE(flag_peer_dev, SGM_ERR_INVALID, "sgm_p2p_signal: null flag");
  p2p_signal_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(flag_peer_dev, value);
  SGM_CUDA_CHECK(cudaGetLastError());
  return SGM_OK;
}

extern "C" int32_t sgm_p2p_wait(const uint32_t* flag_dev, uint32_t value, int32_t* timeout_flag_dev, double timeout_s,
                                void* stream) {
  SGM_REQUIRE(flag_dev, SGM_ERR_INVALID, "sgm_p2p_wait: null flag");
  int dev = 0, khz = 0;
  SGM_CUDA_CHECK(cudaGetDevice(&dev));
  SGM_CUDA_CHECK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));
  const long long max_cycles = (long long)((timeout_s > 0 ? timeout_s : 30.0) * 1000.0 * (khz > 0 ? khz : 1900000));
  p2p_wait_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(flag_dev, value, timeout_flag_dev, max_cycles);
  SGM_CUDA_CHECK(cudaGetLastError());
  return SGM_OK;
}
