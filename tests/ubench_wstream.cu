// Micro-benchmark (GPU box only): L2 -> shared-memory streaming rate of filter banks that EVERY CTA reads
// (cp.async.bulk ring, one CTA per SM), as a function of the stage size, the ring depth, the size of the bank and of
// whether the CTAs walk the bank in lock step, with a per-CTA phase offset, or own private banks.
// Question behind it: is a persistent conv kernel whose CTAs all stream the same weights bound by the
// L2 slices that hold them?
// Build + run:  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I segmantic_b200/csrc
//               tests/ubench_wstream.cu -o /tmp/ubench_wstream && /tmp/ubench_wstream
#include <cstdio>
#include <cuda_runtime.h>

#include "tc_ptx.cuh"

using namespace sgm::tcptx;

struct Cfg {
  int stage_bytes;  // bytes per bulk copy
  int nstages;      // ring depth
  long long bank;   // bytes of the bank a CTA walks (wraps around)
  int mode;         // 0 lock step (all CTAs same addresses), 1 phase offset per CTA, 2 private bank per CTA
  int iters;        // copies per CTA
  int cluster_mc;   // unused (reserved)
};

__global__ void __launch_bounds__(128, 1) wstream(const Cfg c, const uint8_t* __restrict__ w, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[8];
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < c.nstages; ++s) mbar_init(smem_u32(&bars[s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    const long long nst = c.bank / c.stage_bytes;  // stages per bank
    long long pos = c.mode == 1 ? ((long long)blockIdx.x * 37) % nst : 0;
    const uint8_t* base = w + (c.mode == 2 ? (long long)blockIdx.x * c.bank : 0);
    const long long t0 = clock64();
    for (int it = 0; it < c.iters + c.nstages; ++it) {
      const int s = it % c.nstages;
      if (it >= c.nstages) {  // consume: wait for the copy issued nstages ago
        const uint32_t ph = (uint32_t)((it - c.nstages) / c.nstages) & 1u;
        while (!mbar_try_wait(smem_u32(&bars[s]), ph)) {
        }
      }
      if (it < c.iters) {
        mbar_expect_tx(smem_u32(&bars[s]), (uint32_t)c.stage_bytes);
        bulk_g2s(smem_u32(smem + (size_t)s * c.stage_bytes), base + pos * c.stage_bytes, (uint32_t)c.stage_bytes,
                 smem_u32(&bars[s]));
        pos = pos + 1 == nst ? 0 : pos + 1;
      }
    }
    const long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
}

int main() {
  const long long kBuf = 148LL * 8 * 1024 * 1024;
  uint8_t* w;
  long long* out;
  cudaMalloc(&w, kBuf);
  cudaMemset(w, 1, kBuf);
  cudaMalloc(&out, 148 * sizeof(long long));
  cudaFuncSetAttribute(wstream, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  printf("%8s %6s %10s %6s | %10s %12s %12s\n", "stageB", "depth", "bankB", "mode", "B/clk/SM", "chip B/clk", "clk/stage");
  auto run = [&](Cfg c, int grid) {
    for (int rep = 0; rep < 2; ++rep) wstream<<<grid, 128, 200 * 1024>>>(c, w, out);
    long long h[148];
    cudaError_t e = cudaMemcpy(h, out, grid * sizeof(long long), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) {
      printf("CUDA error: %s\n", cudaGetErrorString(e));
      exit(1);
    }
    long long mx = 0;
    for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
    const double bpc = (double)c.stage_bytes * c.iters / (double)mx;
    printf("%8d %6d %10lld %6d | %10.2f %12.1f %12.1f  (grid %d)\n", c.stage_bytes, c.nstages, c.bank, c.mode, bpc,
           bpc * grid, (double)mx / c.iters, grid);
  };
  for (int grid : {1, 148}) {
    for (int stage : {8192, 16384, 32768}) {
      for (int depth : {2, 4}) {
        if ((long long)stage * depth > 190 * 1024) continue;
        for (long long bank : {1LL << 20, 4LL << 20}) {
          for (int mode : {0, 1, 2}) run(Cfg{stage, depth, bank, mode, 2000, 0}, grid);
        }
      }
    }
  }
  // small hot bank (one K-block group re-read by every CTA, e.g. resident-style reloads)
  for (int mode : {0, 1}) run(Cfg{16384, 4, 64 << 10, mode, 2000, 0}, 148);
  return 0;
}
