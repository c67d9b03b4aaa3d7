// Micro-benchmark (GPU box only): does tcgen05.ld traffic of epilogue warps slow down concurrent SS-mode
// tcgen05.mma issue?  Warp 8 issues MMAs (M=128, N, K=16), warps 0..nw-1 stream TMEM columns with tcgen05.ld.
// nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I segmantic_b200/csrc tests/ubench_tmem.cu -o tests/ubench_tmem.bin
#include <cstdio>
#include <cuda_runtime.h>

#include "tc_ptx.cuh"

using namespace sgm::tcptx;

struct Cfg {
  int N, nmma, nw, nld, ldcols, spin, sleep_ns, commit_every, vary_b, fence_every, wait_every;  // nld loads of `ldcols` (16 or 32) columns per epilogue warp
};

__global__ void __launch_bounds__(288, 1) ubench(const Cfg c, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint64_t bar2;
  __shared__ uint64_t bar3[8];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 200 * 1024 / 16; i += 288) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u);
  if (tid == 0) {
    mbar_init(smem_u32(&bar), 1);
    mbar_init(smem_u32(&bar2), 1);
    for (int i = 0; i < 8; ++i) mbar_init(smem_u32(&bar3[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 8) {
    if (elect_one() && c.nmma > 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(c.N >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t a16 = smem_u32(smem) >> 4;
      const uint32_t b16 = (smem_u32(smem) + 160 * 1024) >> 4;
      const uint64_t bdesc = make_desc(b16, (uint32_t)c.N, 8);
      const long long t0 = clock64();
      uint32_t slot = 0;
      for (int g = 0; g < c.nmma / 9; ++g) {
        const uint32_t d = tmem + slot * 128;
        uint32_t a_lo = (a16 & 0x3FFFu) | (2048u << 16);
        const uint32_t a_hi = 8u | (1u << 14);
#pragma unroll
        for (int j = 0; j < 9; ++j, a_lo += 17)
          tc_mma(d, ((uint64_t)a_hi << 32) | a_lo, bdesc + (c.vary_b ? (uint64_t)(j * c.N * 2) : 0ull), idesc, j ? 1u : 0u);
        slot = (slot + 1) & 1;
        if (c.commit_every) tc_commit(smem_u32(&bar3[g & 7]));
        if (c.wait_every) { while (!mbar_try_wait(smem_u32(&bar3[7]), 1u)) {} }
        if (c.fence_every) tc_fence_after();
      }
      const long long t1 = clock64();
      tc_commit(smem_u32(&bar));
      while (!mbar_try_wait(smem_u32(&bar), 0u)) {
      }
      const long long t2 = clock64();
      out[0] = t1 - t0, out[1] = t2 - t0;
      mbar_arrive(smem_u32(&bar2));
    }
  } else if (warp >= c.nw && warp < c.nw + c.spin && c.nmma > 0) {
    while (!mbar_try_wait(smem_u32(&bar2), 0u)) {
      if (c.sleep_ns) __nanosleep(c.sleep_ns);
    }
  } else if (warp < c.nw) {
    const uint32_t tl = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 256;  // columns 256.. (not the accumulators)
    uint32_t acc = 0;
    const long long t0 = clock64();
    for (int i = 0; i < c.nld; ++i) {
      uint32_t v[32];
      if (c.ldcols == 32) {
        tc_ld16(tl + (i & 7) * 32, v);
        tc_ld16(tl + (i & 7) * 32 + 16, v + 16);
        acc += v[0] + v[31];
      } else {
        tc_ld16(tl + (i & 15) * 16, v);
        acc += v[0] + v[15];
      }
    }
    const long long t1 = clock64();
    if (lane == 0) out[2 + warp] = t1 - t0;
    if (acc == 0x12345678u) out[15] = acc;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
  long long* out;
  cudaMalloc(&out, 16 * sizeof(long long));
  cudaFuncSetAttribute(ubench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  printf("%4s %6s %3s %6s %6s | %9s %9s | %12s %14s\n", "N", "nmma", "nw", "nld", "cols", "issue/mma", "done/mma", "clk/ld(warp0)", "B/clk all warps");
  auto run = [&](Cfg c) {
    cudaMemset(out, 0, 16 * sizeof(long long));
    ubench<<<1, 288, 200 * 1024>>>(c, out);
    long long h[16];
    cudaError_t e = cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) {
      printf("CUDA error: %s\n", cudaGetErrorString(e));
      exit(1);
    }
    const double ldclk = c.nw && c.nld ? (double)h[2] / c.nld : 0.0;
    const double bpc = ldclk > 0 ? c.nw * 32.0 * c.ldcols * 4.0 / ldclk : 0.0;
    printf("%4d %6d %3d %6d %6d spin%d/%d commit%d varyB%d fence%d wait%d | %9.1f %9.1f | %12.1f %14.1f\n", c.N, c.nmma, c.nw, c.nld, c.ldcols, c.spin, c.sleep_ns, c.commit_every, c.vary_b, c.fence_every, c.wait_every,
           c.nmma ? (double)h[0] / c.nmma : 0.0, c.nmma ? (double)h[1] / c.nmma : 0.0, ldclk, bpc);
  };
  for (int N : {32, 48, 96}) {
    run(Cfg{N, 2700, 0, 0, 16});
    for (int nw : {1, 4, 8}) {
      run(Cfg{N, 0, nw, 4000, 16});
      run(Cfg{N, 2700, nw, 4000, 16});
      run(Cfg{N, 2700, nw, 4000, 32});
    }
    run(Cfg{N, 2700, 0, 0, 16, 0, 0, 1, 0});
    run(Cfg{N, 2700, 0, 0, 16, 0, 0, 0, 1});
    run(Cfg{N, 2700, 0, 0, 16, 0, 0, 1, 1});
    run(Cfg{N, 2700, 0, 0, 16, 0, 0, 1, 1, 1, 0});
    run(Cfg{N, 2700, 0, 0, 16, 0, 0, 1, 1, 0, 1});
    run(Cfg{N, 2700, 0, 0, 16, 0, 0, 1, 1, 1, 1});
    for (int spin : {8}) {
      run(Cfg{N, 2700, 0, 0, 16, spin, 0});
      run(Cfg{N, 2700, 0, 0, 16, spin, 32});
    }
  }
  return 0;
}
