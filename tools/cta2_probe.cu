// Probe of the 2-CTA (cta_group::2) tcgen05 path on sm_100a, with the no-swizzle K-major operand layouts this
// repo uses: a CTA pair multiplies A[256 x K] (128 rows per CTA) by B[N x K] (N/2 rows per CTA) into TMEM of both
// CTAs, commit multicast to both, each CTA reads its 128 x N block back.  Checks the result against the host and
// times M=256,N=128,K=16 issue rate.   GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/cta2_probe tools/cta2_probe.cu && timeout 60 /tmp/cta2_probe
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include "../audio-watermarking-deep-learning-watermarks-for-authenticating-speech_b200/csrc/wm_tc.cuh"
using namespace wm::tc;

__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mma_bf16_2cta(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit_2cta(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask) : "memory");
}

// A: [2][128][K] bf16 row-major (CTA r owns rows 128r..), B: [N][K] bf16 (CTA r owns rows N/2*r ..), D: [256][N] fp32
template <int N, int K>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
    probe(const __nv_bfloat16 *A, const __nv_bfloat16 *B, float *D, long long *cyc, int reps) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const uint32_t rank = cluster_rank();
  const uint32_t sb = smem_u32(smem);
  constexpr int A_B = 128 * K * 2, KC = K / 8;
  // K-major no-swizzle: [k chunk][row][8 elems]; LBO = rows*16, SBO = 128
  for (int e = threadIdx.x; e < 128 * K; e += blockDim.x) {
    const int r = e / K, k = e % K;
    reinterpret_cast<__nv_bfloat16 *>(smem)[(k / 8) * 128 * 8 + r * 8 + (k % 8)] = A[((size_t)rank * 128 + r) * K + k];
  }
  for (int e = threadIdx.x; e < (N / 2) * K; e += blockDim.x) {
    const int r = e / K, k = e % K;
    reinterpret_cast<__nv_bfloat16 *>(smem + A_B)[(k / 8) * (N / 2) * 8 + r * 8 + (k % 8)] = B[((size_t)rank * (N / 2) + r) * K + k];
  }
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "n"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (rank == 0 && threadIdx.x < 32) {
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc(256, N);
      long long t0 = clock64();
      for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int kk = 0; kk < K / 16; ++kk) {
          const uint64_t ad = smem_desc(sb + (2 * kk) * 128 * 16, 128 * 16, 128);
          const uint64_t bd = smem_desc(sb + A_B + (2 * kk) * (N / 2) * 16, (N / 2) * 16, 128);
          mma_bf16_2cta(tmem, ad, bd, idesc, (r | kk) ? 1u : 0u);
        }
      }
      long long t1 = clock64();
      commit_2cta(smem_u32(&bar), 3);
      if (cyc) { cyc[0] = t1 - t0; }
    }
    __syncwarp();
  }
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  {  // each warp reads its 32 lanes
    const int q = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int c0 = 0; c0 < N; c0 += 16) {
      float v[16];
      tmem_ld16(tmem + c0 + ((uint32_t)(q * 32) << 16), v);
      tmem_ld_wait();
      for (int c = 0; c < 16; ++c) D[((size_t)rank * 128 + q * 32 + lane) * N + c0 + c] = v[c];
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (threadIdx.x < 32) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(256) : "memory");
  }
}

template <int N, int K>
int run(int reps) {
  std::vector<__nv_bfloat16> hA(256 * K), hB(N * K);
  std::vector<float> fA(256 * K), fB(N * K);
  srand(1);
  for (size_t i = 0; i < hA.size(); ++i) { float v = (rand() % 17 - 8) / 8.0f; hA[i] = __float2bfloat16(v); fA[i] = v; }
  for (size_t i = 0; i < hB.size(); ++i) { float v = (rand() % 13 - 6) / 4.0f; hB[i] = __float2bfloat16(v); fB[i] = v; }
  __nv_bfloat16 *dA, *dB; float *dD; long long *dc;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, 256 * N * 4); cudaMalloc(&dc, 16);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, 256 * N * 4);
  const int smem = 128 * K * 2 + (N / 2) * K * 2 + 1024;
  cudaFuncSetAttribute(probe<N, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe<N, K><<<2, 128, smem>>>(dA, dB, dD, dc, reps);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("N=%d K=%d: CUDA error %s\n", N, K, cudaGetErrorString(e)); return 1; }
  std::vector<float> hD(256 * N);
  long long cyc = 0;
  cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(&cyc, dc, 8, cudaMemcpyDeviceToHost);
  double maxerr = 0;
  for (int m = 0; m < 256; ++m)
    for (int n = 0; n < N; ++n) {
      double s = 0;
      for (int k = 0; k < K; ++k) s += (double)fA[m * K + k] * fB[n * K + k];
      s *= reps;
      double d = fabs(s - hD[m * N + n]);
      if (d > maxerr) maxerr = d;
    }
  printf("N=%d K=%d reps=%d: max abs err %.3g, %.1f cycles per 256x%dx16 MMA\n", N, K, reps, maxerr,
         (double)cyc / (reps * (K / 16)), N);
  return maxerr < 1e-3 * reps ? 0 : 2;
}

int main() {
  int rc = 0;
  rc |= run<64, 16>(1);
  rc |= run<128, 16>(1);
  rc |= run<128, 64>(1);
  rc |= run<128, 64>(512);
  rc |= run<64, 64>(512);
  printf(rc ? "FAILED\n" : "OK\n");
  return rc;
}
