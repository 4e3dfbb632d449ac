// Micro-benchmark of tcgen05.mma issue/execution rate on sm_100a for the operand layouts this
// repo uses (no-swizzle K-major smem tiles, 128B-swizzled tiles, A in TMEM).  GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mma_bench tools/mma_bench.cu && /tmp/mma_bench
#include <cstdio>
#include <cuda_runtime.h>
#include "../audio-watermarking-deep-learning-watermarks-for-authenticating-speech_b200/csrc/wm_tc.cuh"
using namespace wm::tc;

__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

// N: MMA N; NACC: accumulators cycled; A_TMEM: A operand from tensor memory; SWZ: 128B-swizzle descriptors
// ELECT: issue from a converged warp under elect.sync instead of a lane-0 branch
template <int N, int NACC, int A_TMEM, int SWZ, int ELECT, int M = 128>
__global__ void __launch_bounds__(128, 1) bench(int reps, long long *out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const uint32_t sb = smem_u32(smem);
  for (int i = threadIdx.x; i < 160 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (threadIdx.x < 32) {
    constexpr uint32_t idesc = make_idesc(M, N);
    uint64_t ad[4], bd[4];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      ad[kk] = SWZ ? desc_sw128(sb + kk * 32) : smem_desc(sb + (2 * kk) * 2080, 2080, 128);
      bd[kk] = SWZ ? desc_sw128(sb + 65536 + kk * 32) : smem_desc(sb + 65536 + (2 * kk) * N * 16, N * 16, 128);
    }
    uint32_t elected = 0;
    if (ELECT) {
      asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(elected));
    }
    const bool issuer = ELECT ? (elected != 0) : (threadIdx.x == 0);
    long long t0 = clock64();
#pragma unroll 1
    for (int r = 0; r < reps; r += 16) {
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const int kk = u & 3;
        const uint32_t d = tmem + 256 + (u % NACC) * N;
        if (issuer) {
          if (A_TMEM) mma_bf16_ts(d, tmem + kk * 8, bd[kk], idesc, 1);
          else mma_bf16(d, ad[kk], bd[kk], idesc, 1);
        }
      }
    }
    long long t1 = clock64();
    if (issuer) {
      tc_commit(smem_u32(&bar));
      mbar_wait(smem_u32(&bar), 0);
      long long t2 = clock64();
      if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
  }
}

template <int N, int NACC, int A_TMEM, int SWZ, int ELECT, int M = 128>
void run(long long *out, int grid) {
  const int reps = 4096;
  cudaFuncSetAttribute(bench<N, NACC, A_TMEM, SWZ, ELECT, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  bench<N, NACC, A_TMEM, SWZ, ELECT, M><<<grid, 128, 160 * 1024>>>(reps, out);
  long long h[2];
  cudaError_t e = cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
  printf("%d,%d,%d,%d,%d,%d,%.1f,%.1f%s\n", N, NACC, A_TMEM, SWZ, ELECT, grid, (double)h[0] / reps, (double)h[1] / reps,
         M == 64 ? ",M=64" : "");
}

int main() {
  long long *out;
  cudaMalloc(&out, 16);
  printf("N,nacc,a_tmem,swizzle128,elect,grid,issue_cyc_per_mma,total_cyc_per_mma\n");
  for (int grid : {1, 148}) {
    run<32, 1, 0, 0, 0>(out, grid);  run<64, 1, 0, 0, 0>(out, grid);  run<128, 1, 0, 0, 0>(out, grid);  run<256, 1, 0, 0, 0>(out, grid);
    run<64, 2, 0, 0, 0>(out, grid);  run<128, 2, 0, 0, 0>(out, grid);
    run<32, 1, 0, 1, 0>(out, grid);  run<64, 1, 0, 1, 0>(out, grid);  run<128, 1, 0, 1, 0>(out, grid);  run<256, 1, 0, 1, 0>(out, grid);
    run<128, 2, 0, 1, 0>(out, grid);
    run<32, 1, 1, 0, 0>(out, grid);  run<64, 1, 1, 0, 0>(out, grid);  run<128, 1, 1, 0, 0>(out, grid);  run<256, 1, 1, 0, 0>(out, grid);
    run<64, 2, 1, 0, 0>(out, grid);  run<64, 4, 1, 0, 0>(out, grid);
    run<64, 1, 1, 1, 0>(out, grid);  run<128, 1, 1, 1, 0>(out, grid);
    run<64, 1, 0, 0, 1>(out, grid);  run<128, 1, 0, 0, 1>(out, grid);  run<128, 1, 0, 1, 1>(out, grid);  run<64, 1, 1, 0, 1>(out, grid);
    // small-N shapes of the LSTM (A = weights in TMEM)
    run<32, 1, 1, 0, 1>(out, grid);  run<16, 1, 1, 0, 1>(out, grid);  run<32, 2, 1, 0, 1>(out, grid);  run<16, 4, 1, 0, 1>(out, grid);
    run<32, 1, 0, 0, 1>(out, grid);  run<16, 1, 0, 0, 1>(out, grid);  run<256, 1, 0, 0, 1>(out, grid);  run<256, 1, 1, 0, 1>(out, grid);
    // M = 64 (operand-role swap study: weights as the A operand, activations as B with N = 256)
    run<256, 1, 1, 0, 1, 64>(out, grid);  run<256, 1, 0, 0, 1, 64>(out, grid);  run<128, 1, 1, 0, 1, 64>(out, grid);
    run<128, 2, 1, 0, 1>(out, grid);
  }
  return 0;
}
