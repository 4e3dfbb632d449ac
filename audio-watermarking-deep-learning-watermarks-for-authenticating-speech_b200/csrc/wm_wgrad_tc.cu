// Weight gradient of Conv1d(64,64,K,padding=K/2) on the tensor cores (sm_100a), for the training step
// (py/main16.py:277; the convolutions of ResBlock :116,119 and the ConvTranspose1d :144 in its convolution form):
//     dW[j][ci][co] = sum over clips b and time t of  x[b][t + j - K/2][ci] * dy[b][t][co]
// As a GEMM the contraction runs over TIME, so both operands are "MN-major" -- and a plane of the library's planar
// activation format ([t][8 channels x bf16], 16-byte rows) IS the tcgen05 no-swizzle MN-major canonical layout:
// eight consecutive time rows x 16 bytes form one core matrix, SBO = plane pitch steps to the next 8 channels,
// LBO = 128 bytes to the next 8 time rows.  With M = 128 the A descriptor walks all 16 planes of x, i.e.
// [x_hi ; x_lo] (64 channels each), and with N = 128 the B descriptor walks [dy_hi | dy_lo]: ONE 128x128x16 MMA per
// tap and 16 time steps produces all four bf16 partial products, which the epilogue adds
// (D[ci][co] + D[ci][64+co] + D[64+ci][co] + D[64+ci][64+co]).  Tap j is the same x tile with the descriptor start
// address moved by j rows.  Each CTA accumulates its share of the (clip, time-tile) list in tensor memory for the
// whole kernel (no per-tile epilogue) and writes one partial [tap][128 lanes][64]; a second kernel adds the
// partials in a fixed order (deterministic) and folds the hi/lo lane halves.
//
// Shared-memory traffic is balanced with the tensor pipe (8 KB of operands per 64-cycle MMA); the kernel is bound by
// HBM: it reads x and dy once in the planar format (4 B per element each).
// TMEM: taps x 128 columns, so K = 7 runs as two launches (taps 0..3, 4..6).
#include <cuda_bf16.h>

#include "wm_common.h"
#include "wm_tc.cuh"

namespace wm {

using namespace tc;

namespace {

constexpr int PAD = WM_PLANAR_PAD;
constexpr int TILE = 128;                 // time rows per tile = 8 MMA K-steps
constexpr int MAXT = 4;                   // taps per launch (4 x 128 TMEM columns)
constexpr int XROWS = TILE + MAXT - 1;    // rows of an x tile buffer
constexpr int X_PLANE_B = XROWS * 16, Y_PLANE_B = TILE * 16;
constexpr int X_TILE_B = 16 * X_PLANE_B, Y_TILE_B = 16 * Y_PLANE_B;
constexpr int STAGE_B = X_TILE_B + Y_TILE_B;        // 66 304
constexpr int NSTAGE = 3;
constexpr int OFF_BAR = NSTAGE * STAGE_B;
constexpr int WG_SMEM = OFF_BAR + 128;
constexpr int WG_THREADS = 192;           // warps 0..3 epilogue (one TMEM lane quadrant each), 4 producer, 5 MMA issue
// D fp32, A / B bf16, both MN-major (bits 15, 16), M = 128, N = 128
constexpr uint32_t kIdescMN = make_idesc(128, 128) | (1u << 15) | (1u << 16);

}  // namespace

// xp, dyp: planar [clip][16 planes][T + 2 PAD rows][16 B]; partial: [gridDim.x][ntaps][128][64] fp32
template <int NTAPS>   // taps of this launch (compile-time: a run-time guard around the MMAs makes ptxas wrap each one in a loop)
__global__ void __launch_bounds__(WG_THREADS, 1)
    wgrad_tc_kernel(const uint4 *__restrict__ xp, const uint4 *__restrict__ dyp, float *__restrict__ partial, int B, int T,
                    int tap0, int pad) {
  constexpr int ntaps = NTAPS;
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t s_base = smem_u32(smem);
  const uint32_t bars = s_base + OFF_BAR;
  auto full_bar = [&](int s) { return bars + 8 * s; };
  auto empty_bar = [&](int s) { return bars + 8 * (NSTAGE + s); };
  const uint32_t done_bar = bars + 8 * (2 * NSTAGE);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + OFF_BAR + 8 * (2 * NSTAGE + 1));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntile_t = (T + TILE - 1) / TILE;
  const long long ntiles = (long long)B * ntile_t;
  const size_t RP = (size_t)T + 2 * PAD;
  const int xrows = TILE + ntaps - 1;           // rows of x a tile needs: t0 - pad + tap0 .. + TILE + ntaps - 2

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(done_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 4) {
    // ===== producer: per tile 16 x-plane and 16 dy-plane bulk copies; rows past the end of a clip are zeroed =====
    int i = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++i) {
      const int s = i % NSTAGE;
      const uint32_t ph = (i / NSTAGE) & 1;
      const long long b = tile / ntile_t;
      const int t0 = (int)(tile % ntile_t) * TILE;
      // valid rows: dy rows t0 .. T-1; x rows up to the planes' zero padding (row index T + PAD - 1 of the plane)
      const int yrows_ok = min(TILE, T - t0);
      const int x_first = t0 + PAD - pad + tap0;                       // plane row of the tile's first x row (>= 0)
      const int xrows_ok = min(xrows, (int)RP - x_first);
      mbar_wait_warp(empty_bar(s), ph ^ 1);
      uint8_t *stage = smem + (size_t)s * STAGE_B;
      if (yrows_ok < TILE || xrows_ok < xrows) {   // (last tile of a clip only) zero what the copies will not write
        for (int e = lane; e < 16 * (TILE - yrows_ok); e += 32) {
          const int p = e / (TILE - yrows_ok), r = yrows_ok + e % (TILE - yrows_ok);
          *reinterpret_cast<uint4 *>(stage + X_TILE_B + p * Y_PLANE_B + r * 16) = make_uint4(0, 0, 0, 0);
        }
        for (int e = lane; e < 16 * (xrows - xrows_ok); e += 32) {
          const int p = e / (xrows - xrows_ok), r = xrows_ok + e % (xrows - xrows_ok);
          *reinterpret_cast<uint4 *>(stage + p * X_PLANE_B + r * 16) = make_uint4(0, 0, 0, 0);
        }
        fence_async_smem();
      }
      __syncwarp();
      if (lane == 0) {
        mbar_arrive_expect_tx(full_bar(s), 16u * (uint32_t)(xrows_ok + yrows_ok) * 16u);
#pragma unroll 4
        for (int p = 0; p < 16; ++p) {
          bulk_g2s(s_base + s * STAGE_B + p * X_PLANE_B, xp + ((size_t)(b * 16 + p) * RP + x_first), xrows_ok * 16,
                   full_bar(s));
          bulk_g2s(s_base + s * STAGE_B + X_TILE_B + p * Y_PLANE_B, dyp + ((size_t)(b * 16 + p) * RP + PAD + t0),
                   yrows_ok * 16, full_bar(s));
        }
      }
      __syncwarp();
    }
  } else if (warp == 5) {
    // ===== MMA issuer: everything accumulates into the same [tap] accumulators =====
    const bool issuer = elect_one();
    int i = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++i) {
      const int s = i % NSTAGE;
      const uint32_t ph = (i / NSTAGE) & 1;
      mbar_wait_warp(full_bar(s), ph);
      tc_fence_after();
      if (issuer) {
        // MN-major, no swizzle: LBO = 128 B (next 8 time rows), SBO = plane pitch (next 8 channels)
        const uint64_t a0 = smem_desc(s_base + s * STAGE_B, 128, X_PLANE_B);
        const uint64_t b0 = smem_desc(s_base + s * STAGE_B + X_TILE_B, 128, Y_PLANE_B);
#pragma unroll
        for (int ks = 0; ks < TILE / 16; ++ks) {
#pragma unroll
          for (int j = 0; j < NTAPS; ++j) {
            mma_bf16(tmem + j * 128, a0 + (uint64_t)(((16 * ks + j) * 16) >> 4), b0 + (uint64_t)((16 * ks * 16) >> 4),
                     kIdescMN, (i | ks) != 0 ? 1u : 0u);
          }
        }
        tc_commit(empty_bar(s));
      }
      __syncwarp();
    }
    if (issuer) tc_commit(done_bar);
    __syncwarp();
  } else {
    // ===== epilogue (once): accumulator lanes -> partial[cta][tap][lane][64], hi/lo column halves added =====
    const long long my_tiles = ntiles > blockIdx.x ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    float *dst = partial + ((size_t)blockIdx.x * ntaps * 128 + warp * 32 + lane) * 64;
    if (my_tiles > 0) {
      mbar_wait_warp(done_bar, 0);
      tc_fence_after();
    }
    for (int j = 0; j < ntaps; ++j) {
      float *d = dst + (size_t)j * 128 * 64;
#pragma unroll 1
      for (int c0 = 0; c0 < 64; c0 += 16) {
        float v1[16], v2[16];
        if (my_tiles > 0) {
          tmem_ld16(tmem + j * 128 + c0 + ((uint32_t)(warp * 32) << 16), v1);
          tmem_ld16(tmem + j * 128 + 64 + c0 + ((uint32_t)(warp * 32) << 16), v2);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int c = 0; c < 16; ++c) v1[c] = v2[c] = 0.0f;
        }
#pragma unroll
        for (int c = 0; c < 16; c += 4)
          *reinterpret_cast<float4 *>(d + c0 + c) =
              make_float4(v1[c] + v2[c], v1[c + 1] + v2[c + 1], v1[c + 2] + v2[c + 2], v1[c + 3] + v2[c + 3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
  }
}

// dw[tap0 + j][ci][co] = sum over CTAs (ascending) of partial[cta][j][ci][co] + partial[cta][j][64 + ci][co]
__global__ void __launch_bounds__(256)
    wgrad_reduce_kernel(const float *__restrict__ partial, float *__restrict__ dw, int nparts, int tap0, int ntaps) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= ntaps * 4096) return;
  const int j = e >> 12, ci = (e >> 6) & 63, co = e & 63;
  float acc = 0.0f;
  for (int c = 0; c < nparts; ++c) {
    const float *p = partial + ((size_t)c * ntaps + j) * 128 * 64;
    acc += p[ci * 64 + co] + p[(64 + ci) * 64 + co];
  }
  dw[(size_t)(tap0 + j) * 4096 + ci * 64 + co] = acc;
}

// db[co] = sum over rows of dy[rows][64]: per-block partial sums in a fixed order, then one block adds them (deterministic)
__global__ void __launch_bounds__(256)
    colsum64_partial_kernel(const float *__restrict__ dy, float *__restrict__ part, long long rows, long long rows_per_block) {
  __shared__ float red[4][64];
  const int c = threadIdx.x & 63, r = threadIdx.x >> 6;
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  const long long r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
  float acc = 0.0f;
  for (long long i = r0 + r; i < r1; i += 4) acc += dy[i * 64 + c];
  red[r][c] = acc;
  __syncthreads();
  if (r == 0) part[(size_t)blockIdx.x * 64 + c] = (red[0][c] + red[1][c]) + (red[2][c] + red[3][c]);
}
__global__ void colsum64_final_kernel(const float *__restrict__ part, float *__restrict__ db, int nblocks) {
  const int c = threadIdx.x;
  double acc = 0.0;
  for (int b = 0; b < nblocks; ++b) acc += (double)part[(size_t)b * 64 + c];
  db[c] = (float)acc;
}
constexpr int kColsumBlocks = 1184;
size_t colsum64_scratch_floats() { return (size_t)kColsumBlocks * 64; }
int launch_colsum64(const float *dy, float *db, long long rows, float *scratch, cudaStream_t st) {
  long long per = (rows + kColsumBlocks - 1) / kColsumBlocks;
  if (per < 4) per = 4;
  const int nb = (int)((rows + per - 1) / per);
  colsum64_partial_kernel<<<nb, 256, 0, st>>>(dy, scratch, rows, per);
  WM_CHECK_LAUNCH("colsum64_partial");
  colsum64_final_kernel<<<1, 64, 0, st>>>(scratch, db, nb);
  WM_CHECK_LAUNCH("colsum64_final");
  return 0;
}

size_t wgrad_tc_scratch_floats(int K) {
  const int nt = K < MAXT ? K : MAXT;
  return (size_t)sm_count() * nt * 128 * 64;
}

// xp, dyp planar; dw [K][64][64] tap-major; scratch: wgrad_tc_scratch_floats(K) floats
int launch_wgrad_tc(const void *xp, const void *dyp, float *dw, int B, int T, int K, float *scratch, cudaStream_t st) {
  if (B == 0 || T == 0) return 0;
  static bool attr_set = false;
  if (!attr_set) {
    WM_CHECK_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM));
    WM_CHECK_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM));
    attr_set = true;
  }
  if (K != 3 && K != 7) {
    set_error("wgrad_tc: K must be 3 or 7 (got %d)", K);
    return -1;
  }
  const long long ntiles = (long long)B * ((T + TILE - 1) / TILE);
  const int grid = (int)(ntiles < sm_count() ? ntiles : sm_count());
  for (int tap0 = 0; tap0 < K; tap0 += MAXT) {
    const int nt = K - tap0 < MAXT ? K - tap0 : MAXT;          // 3 (K = 3), or 4 then 3 (K = 7)
    auto kern = nt == 4 ? wgrad_tc_kernel<4> : wgrad_tc_kernel<3>;
    kern<<<grid, WG_THREADS, WG_SMEM, st>>>(reinterpret_cast<const uint4 *>(xp), reinterpret_cast<const uint4 *>(dyp), scratch, B,
                                            T, tap0, K / 2);
    WM_CHECK_LAUNCH("wgrad_tc");
    wgrad_reduce_kernel<<<(nt * 4096 + 255) / 256, 256, 0, st>>>(scratch, dw, grid, tap0, nt);
    WM_CHECK_LAUNCH("wgrad_reduce");
  }
  return 0;
}

}  // namespace wm
