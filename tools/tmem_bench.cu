// TMEM load (tcgen05.ld) bandwidth on sm_100a, alone and under concurrent tcgen05.mma traffic.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tmem_bench tools/tmem_bench.cu && /tmp/tmem_bench
#include <cstdio>
#include <cuda_runtime.h>
#include "../audio-watermarking-deep-learning-watermarks-for-authenticating-speech_b200/csrc/wm_tc.cuh"
using namespace wm::tc;

// NW load warps (multiple of 4) each issue `reps` x (tcgen05.ld.32x32b.x32 + wait); optional MMA warp keeps the pipe busy
template <int WITH_MMA>
__global__ void __launch_bounds__(32 * 17, 1) bench(int nw, int reps, long long *out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t slot;
  __shared__ volatile int stop;
  const uint32_t sb = smem_u32(smem);
  for (int i = threadIdx.x; i < 96 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) stop = 0;
  if (warp == 16) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (warp < nw) {
    const uint32_t ta = tmem + ((uint32_t)((warp & 3) * 32) << 16) + ((warp >> 2) * 32) % 256;
    uint32_t acc = 0;
    long long t0 = clock64();
#pragma unroll 1
    for (int r = 0; r < reps; ++r) {
      uint32_t v[32];
      tmem_ld32(ta, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) acc ^= v[i];
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = acc; }
    __syncwarp();
    if (warp == 0 && (threadIdx.x & 31) == 0) stop = 1;
  } else if (warp == 16 && WITH_MMA) {
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc(128, 128);
      const uint64_t a0 = smem_desc(sb, 2080, 128), b0 = smem_desc(sb + 49152, 2048, 128);
      while (!stop) {
#pragma unroll
        for (int u = 0; u < 8; ++u) mma_bf16(tmem + 256 + (u & 1) * 128, a0 + (u & 3) * 260, b0 + (u & 3) * 256, idesc, 1);
      }
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 16) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
  }
}

int main() {
  long long *out;
  cudaMalloc(&out, 16);
  cudaFuncSetAttribute(bench<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  cudaFuncSetAttribute(bench<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  const int reps = 2000;
  printf("load_warps,with_mma,cycles_per_ld_x32_per_warp,tmem_read_bytes_per_cycle_SM\n");
  for (int with_mma = 0; with_mma < 2; ++with_mma)
    for (int nw : {4, 8, 16}) {
      if (with_mma) bench<1><<<148, 32 * 17, 96 * 1024>>>(nw, reps, out);
      else bench<0><<<148, 32 * 17, 96 * 1024>>>(nw, reps, out);
      long long h[2];
      cudaError_t e = cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      double cyc = (double)h[0] / reps;
      printf("%d,%d,%.1f,%.1f\n", nw, with_mma, cyc, nw * 32 * 32 * 4 / cyc);
    }
  return 0;
}
