// Micro-benchmark (GPU box only): cost of one SS-mode tcgen05.mma (M=128, K=16, bf16) as a function of N, of
// the alignment of the A start address inside the K-major SWIZZLE_NONE layout, and of the accumulator
// rotation.  Build + run:  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I segmantic_b200/csrc
//                          tests/ubench_mma.cu -o /tmp/ubench_mma && /tmp/ubench_mma
// Output feeds DESIGN.md 4.1 (why the brick layouts keep every tap shift 128-byte aligned).
#include <cstdio>
#include <cuda_runtime.h>

#include "tc_ptx.cuh"

using namespace sgm::tcptx;

struct Cfg {
  int N;         // MMA N
  int a_off16;   // A start offset in 16-byte units (0 = 128-byte aligned core matrices)
  int a_step16;  // added to the A start for every MMA (tap walk); 0 = same tile
  int nslots;    // accumulators rotated every `per_slot` MMAs
  int per_slot;
  int nmma;
  int sbo16;     // A: 8-row group stride
  int lbo16;     // A: K-chunk stride
};

__global__ void __launch_bounds__(128, 1) ubench(const Cfg c, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 200 * 1024 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u);
  if (tid == 0) {
    mbar_init(smem_u32(&bar), 1);
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
  if (warp == 0 && elect_one() && (blockIdx.x == 0 || true)) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(c.N >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t a16 = smem_u32(smem) >> 4;
    const uint32_t b16 = (smem_u32(smem) + 160 * 1024) >> 4;
    const uint64_t bdesc = make_desc(b16, (uint32_t)c.N, 8);
    int slot_stride = 32;
    while (slot_stride < c.N) slot_stride <<= 1;
    for (int rep = 0; rep < 2; ++rep) {
      const long long t0 = clock64();
      const int ngroups = c.nmma / c.per_slot;
      uint32_t slot = 0;
      const uint32_t a_lo0 = ((a16 + c.a_off16) & 0x3FFFu) | (((uint32_t)c.lbo16 & 0x3FFFu) << 16);
      const uint32_t a_hi = ((uint32_t)c.sbo16 & 0x3FFFu) | (1u << 14);
      for (int g = 0; g < ngroups; ++g) {
        const uint32_t d = tmem + slot * slot_stride;
        uint32_t a_lo = a_lo0;
        if (c.per_slot == 9) {
#pragma unroll
          for (int j = 0; j < 9; ++j, a_lo += c.a_step16)
            tc_mma(d, ((uint64_t)a_hi << 32) | a_lo, bdesc, idesc, j ? 1u : 0u);
        } else {
#pragma unroll
          for (int j = 0; j < 27; ++j, a_lo += c.a_step16)
            tc_mma(d, ((uint64_t)a_hi << 32) | a_lo, bdesc, idesc, j ? 1u : 0u);
        }
        slot = (slot + 1 == (uint32_t)c.nslots) ? 0 : slot + 1;
      }
      const long long t1 = clock64();
      tc_commit(smem_u32(&bar));
      while (!mbar_try_wait(smem_u32(&bar), (uint32_t)rep & 1u)) {
      }
      const long long t2 = clock64();
      if (blockIdx.x == 0) {
        out[2 * rep] = t1 - t0;
        out[2 * rep + 1] = t2 - t0;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
  long long* out;
  cudaMalloc(&out, 4 * sizeof(long long));
  cudaFuncSetAttribute(ubench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int Ns[] = {16, 32, 48, 96, 144, 256};
  printf("%4s %6s %6s %6s %6s %5s %5s | %9s %9s\n", "N", "aoff", "astep", "slots", "per", "sbo", "lbo", "issue/mma", "done/mma");
  int grid = 1;
  auto run = [&](Cfg c) {
    ubench<<<grid, 128, 200 * 1024>>>(c, out);
    long long h[4];
    cudaError_t e = cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) {
      printf("CUDA error: %s\n", cudaGetErrorString(e));
      exit(1);
    }
    printf("%4d %6d %6d %6d %6d %5d %5d | %9.1f %9.1f\n", c.N, c.a_off16, c.a_step16, c.nslots, c.per_slot, c.sbo16, c.lbo16,
           (double)h[2] / c.nmma, (double)h[3] / c.nmma);
  };
  for (int N : Ns) {
    for (int off : {0, 1, 4}) run(Cfg{N, off, 0, 1, 27, 270, 8, 2048});
    run(Cfg{N, 0, 8, 1, 27, 270, 8, 2048});    // aligned tap walk
    run(Cfg{N, 0, 17, 1, 27, 270, 8, 2048});   // unaligned tap walk
    run(Cfg{N, 0, 8, N > 128 ? 2 : 4, 9, 270, 8, 2048});     // aligned walk, rotating accumulators, 9 MMAs each
    run(Cfg{N, 0, 17, N > 128 ? 2 : 4, 9, 270, 8, 2048});
  }
  grid = 148;
  printf("-- all 148 SMs busy (block 0 reports)\n");
  for (int N : {32, 48, 96}) run(Cfg{N, 0, 17, 4, 9, 2700, 8, 2048});
  grid = 1;
  // SBO variants (rows of a tile strided: 8-row groups 256 B / 384 B apart)
  for (int sbo : {8, 16, 24, 32}) run(Cfg{48, 0, 8, 4, 9, 270, sbo, 2048});
  for (int sbo : {8, 16, 24, 32}) run(Cfg{16, 0, 8, 4, 27, 270, sbo, 2048});
  // LBO variants
  for (int lbo : {128, 1024, 2048, 3200}) run(Cfg{48, 0, 8, 4, 9, 270, 8, lbo});
  return 0;
}
