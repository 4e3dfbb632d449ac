// A whole ResBlock (py/main16.py:112-125, eval BatchNorm folded) in ONE tcgen05 kernel:
//     y = relu( x + conv2( relu( conv1(x) + b1 ) ) + b2 )
// The intermediate activation never leaves the SM: conv1's accumulator (TMEM) is bias/ReLU'd,
// split into bf16 hi/lo and written straight into a shared-memory tile that is conv2's A
// operand; the residual is read from the x tile that is already in shared memory.  Per output
// tile (126 time steps x 64 channels) HBM sees one x tile read (33 KB) and one y tile write
// (32 KB) instead of 3 reads + 2 writes for the two-kernel form.
//
// Tile geometry (time axis):  output rows  t0 .. t0+125
//                             intermediate t0-1 .. t0+126  (128 rows = the MMA's M)
//                             x tile       t0-2 .. t0+127  (130 rows)
// Operand formats as in wm_conv_tc.cu (planar bf16 hi/lo planes, no-swizzle K-major, tap shift
// = 16-byte start-address shift, B = [W_hi | W_lo] so N = 128 and all four partial products land
// in one accumulator).
//
// Warps: 0 producer (bulk copies), 1 MMA issuer + TMEM owner, 2..5 epilogue.  TMEM: conv1 and
// conv2 accumulators double buffered (4 x 128 columns).  Shared memory: both weight images
// (96 KB), 3 x-tile stages, 1 intermediate tile.  The MMA warp issues conv1 of tile i+1 before
// conv2 of tile i, so the tensor pipe works on the next tile while the epilogue warps turn tile
// i's conv1 accumulator into conv2's operand.
#include <cuda_bf16.h>

#include "wm_common.h"
#include "wm_tc.cuh"

namespace wm {

using namespace tc;

namespace {

constexpr int PAD = WM_PLANAR_PAD;
constexpr int TO = 126;                    // output rows per tile
constexpr int XROWS = 130;
constexpr int PLANE_B = XROWS * 16;        // bytes of one plane of a tile
constexpr int TILE_B = 16 * PLANE_B;       // 33 280
constexpr int W_TAP_B = 8 * 128 * 16;
constexpr int W_IMG_B = 3 * W_TAP_B;       // 49 152
constexpr int NST = 3;
constexpr int OFF_W = 0;
constexpr int OFF_X = OFF_W + 2 * W_IMG_B;
constexpr int OFF_U = OFF_X + NST * TILE_B;
constexpr int OFF_BIAS = OFF_U + TILE_B;
constexpr int OFF_BAR = OFF_BIAS + 512;
constexpr int RB_SMEM = OFF_BAR + 160;
constexpr uint32_t kIdesc = make_idesc(128, 128);
static_assert(RB_SMEM <= 232448, "shared memory budget");

}  // namespace

__global__ void __launch_bounds__(192, 1)
    resblock_tc_kernel(const uint4 *__restrict__ x, const uint4 *__restrict__ w_img, const float *__restrict__ b1,
                       const float *__restrict__ b2, uint4 *__restrict__ y, float *__restrict__ y32, int B, int T) {
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t s_base = smem_u32(smem);
  const uint32_t w_smem = s_base + OFF_W, x_smem = s_base + OFF_X, u_smem = s_base + OFF_U;
  const uint32_t bars = s_base + OFF_BAR;
  // barriers: full[3] 0..2, empty[3] 3..5, wbar 6, d1_full[2] 7,8, d1_empty[2] 9,10, d2_full[2] 11,12,
  //           d2_empty[2] 13,14, u_full 15 (tmem slot behind them)
  auto bar = [&](int i) { return bars + 8 * i; };
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + OFF_BAR + 8 * 17);
  float *bias_s = reinterpret_cast<float *>(smem + OFF_BIAS);   // [0..63] b1, [64..127] b2
  const uint32_t u_empty = bars + 8 * 16;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntile_t = (T + TO - 1) / TO;
  const long long ntiles = (long long)B * ntile_t;
  const size_t RP = (size_t)T + 2 * PAD;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NST; ++s) { mbar_init(bar(s), 1); mbar_init(bar(3 + s), 128); }
    mbar_init(bar(6), 1);
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar(7 + a), 1); mbar_init(bar(9 + a), 128);
      mbar_init(bar(11 + a), 1); mbar_init(bar(13 + a), 128);
    }
    mbar_init(bar(15), 128);
    mbar_init(u_empty, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x >= 64) {
    const int i = threadIdx.x - 64;
    bias_s[i] = i < 64 ? b1[i] : b2[i - 64];
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ===== producer =====
    if (lane == 0) {
      mbar_arrive_expect_tx(bar(6), 2 * W_IMG_B);
      for (int j = 0; j < 6; ++j)
        bulk_g2s(w_smem + j * W_TAP_B, reinterpret_cast<const uint8_t *>(w_img) + (size_t)j * W_TAP_B, W_TAP_B, bar(6));
      int i = 0;
      for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++i) {
        const int s = i % NST;
        mbar_wait(bar(3 + s), ((i / NST) & 1) ^ 1);
        mbar_arrive_expect_tx(bar(s), TILE_B);
        const long long b = tile / ntile_t;
        const int t0 = (int)(tile % ntile_t) * TO;
        const size_t row0 = (size_t)(t0 + PAD - 2);
#pragma unroll 4
        for (int p = 0; p < 16; ++p)
          bulk_g2s(x_smem + s * TILE_B + p * PLANE_B, x + ((size_t)(b * 16 + p) * RP + row0), PLANE_B, bar(s));
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      auto conv = [&](uint32_t a_tile, uint32_t w_img_s, uint32_t d_tmem) {
        uint32_t accum = 0;
#pragma unroll 1
        for (int j = 0; j < 3; ++j) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint64_t bdesc = smem_desc(w_img_s + j * W_TAP_B + (2 * kk) * 2048, 2048, 128);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              mma_bf16(d_tmem, smem_desc(a_tile + (half * 8 + 2 * kk) * PLANE_B + j * 16, PLANE_B, 128), bdesc,
                       kIdesc, accum);
              accum = 1;
            }
          }
        }
      };
      const long long my_tiles = ntiles > blockIdx.x ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
      mbar_wait(bar(6), 0);
      if (my_tiles > 0) {
        mbar_wait(bar(0), 0);
        tc_fence_after();
        conv(x_smem, w_smem, tmem);
        tc_commit(bar(7));
      }
      for (long long i = 0; i < my_tiles; ++i) {
        const int a = (int)(i & 1);
        if (i + 1 < my_tiles) {  // conv1 of the next tile
          const long long n = i + 1;
          const int s = (int)(n % NST), an = (int)(n & 1);
          mbar_wait(bar(s), (uint32_t)((n / NST) & 1));
          if (n >= 2) mbar_wait(bar(9 + an), (uint32_t)(((n >> 1) - 1) & 1));
          tc_fence_after();
          conv(x_smem + s * TILE_B, w_smem, tmem + an * 128);
          tc_commit(bar(7 + an));
        }
        // conv2 of this tile
        mbar_wait(bar(15), (uint32_t)(i & 1));
        if (i >= 2) mbar_wait(bar(13 + a), (uint32_t)(((i >> 1) - 1) & 1));
        tc_fence_after();
        conv(u_smem, w_smem + W_IMG_B, tmem + 256 + a * 128);
        tc_commit(bar(11 + a));
        tc_commit(u_empty);
      }
    }
    __syncwarp();
  } else {
    // ===== epilogue =====
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const long long my_tiles = ntiles > blockIdx.x ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    auto epi2 = [&](long long i) {
      const long long tile = blockIdx.x + i * gridDim.x;
      const long long b = tile / ntile_t;
      const int t0 = (int)(tile % ntile_t) * TO;
      const int a = (int)(i & 1), s = (int)(i % NST);
      const int t = t0 + row;
      const bool live = row < TO && t < T;
      if (y != nullptr && q == 0 && lane < 2 * PAD) {  // the planes' zero padding rows
        const bool head = lane < PAD;
        if (head ? (t0 == 0) : (t0 + TO >= T)) {
          const size_t zr = head ? (size_t)lane : (size_t)T + lane;
          for (int pl = 0; pl < 16; ++pl) y[((size_t)(b * 16 + pl)) * RP + zr] = make_uint4(0, 0, 0, 0);
        }
      }
      mbar_wait(bar(11 + a), (uint32_t)((i >> 1) & 1));
      tc_fence_after();
      const uint32_t taddr = tmem + 256 + a * 128 + lane_off;
      const uint8_t *xs = smem + OFF_X + s * TILE_B + (row + 2) * 16;   // residual = x rows 2..
      const size_t prow = (size_t)t + PAD;
#pragma unroll 1
      for (int p = 0; p < 4; ++p) {
        float v1[16], v2[16];
        tmem_ld16(taddr + p * 16, v1);
        tmem_ld16(taddr + 64 + p * 16, v2);
        tmem_ld_wait();
        if (p == 3) {
          tc_fence_before();
          mbar_arrive(bar(13 + a));
        }
        float o[16];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int ch = p * 2 + h;
          const uint4 rh = *reinterpret_cast<const uint4 *>(xs + ch * PLANE_B);
          const uint4 rl = *reinterpret_cast<const uint4 *>(xs + (8 + ch) * PLANE_B);
          float r[8];
          join8(rh, rl, r);
#pragma unroll
          for (int c = 0; c < 8; ++c)
            o[h * 8 + c] = fmaxf(v1[h * 8 + c] + v2[h * 8 + c] + bias_s[64 + p * 16 + h * 8 + c] + r[c], 0.0f);
        }
        if (live) {
          if (y != nullptr) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int ch = p * 2 + h;
              uint4 hi, lo;
              split8(o + h * 8, hi, lo);
              y[((size_t)(b * 16 + ch)) * RP + prow] = hi;
              y[((size_t)(b * 16 + 8 + ch)) * RP + prow] = lo;
            }
          }
          if (y32 != nullptr) {
            float4 *dst = reinterpret_cast<float4 *>(y32 + ((size_t)b * T + t) * 64 + p * 16);
#pragma unroll
            for (int c = 0; c < 4; ++c) dst[c] = make_float4(o[4 * c], o[4 * c + 1], o[4 * c + 2], o[4 * c + 3]);
          }
        }
      }
      mbar_arrive(bar(3 + s));   // the x stage (conv1 operand + residual) is free
    };

    for (long long i = 0; i < my_tiles; ++i) {
      // ---- conv1 accumulator -> intermediate tile (conv2's A operand) ----
      const long long tile = blockIdx.x + i * gridDim.x;
      const int t0 = (int)(tile % ntile_t) * TO;
      const int a = (int)(i & 1);
      const int tu = t0 - 1 + row;
      const bool inside = tu >= 0 && tu < T;     // conv2 zero-pads the intermediate feature map
      mbar_wait(bar(7 + a), (uint32_t)((i >> 1) & 1));
      if (i >= 1) mbar_wait(u_empty, (uint32_t)((i - 1) & 1));
      tc_fence_after();
      const uint32_t taddr = tmem + a * 128 + lane_off;
      uint8_t *us = smem + OFF_U + row * 16;
#pragma unroll 1
      for (int p = 0; p < 4; ++p) {
        float v1[16], v2[16];
        tmem_ld16(taddr + p * 16, v1);
        tmem_ld16(taddr + 64 + p * 16, v2);
        tmem_ld_wait();
        if (p == 3) {
          tc_fence_before();
          mbar_arrive(bar(9 + a));
        }
        float o[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) o[c] = inside ? fmaxf(v1[c] + v2[c] + bias_s[p * 16 + c], 0.0f) : 0.0f;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int ch = p * 2 + h;
          uint4 hi, lo;
          split8(o + h * 8, hi, lo);
          *reinterpret_cast<uint4 *>(us + ch * PLANE_B) = hi;
          *reinterpret_cast<uint4 *>(us + (8 + ch) * PLANE_B) = lo;
        }
      }
      fence_async_smem();
      mbar_arrive(bar(15));
      // ---- conv2 accumulator of the previous tile -> y ----
      if (i >= 1) epi2(i - 1);
    }
    if (my_tiles > 0) epi2(my_tiles - 1);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
  }
}

int launch_resblock_tc(const void *x, const void *w_img, const float *b1, const float *b2, void *y, float *y32, int B,
                       int T, cudaStream_t st) {
  if (B == 0 || T == 0) return 0;
  static bool attr_set = false;
  if (!attr_set) {
    WM_CHECK_CUDA(cudaFuncSetAttribute(resblock_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RB_SMEM));
    attr_set = true;
  }
  long long ntiles = (long long)B * ((T + TO - 1) / TO);
  int grid = (int)(ntiles < sm_count() ? ntiles : sm_count());
  resblock_tc_kernel<<<grid, 192, RB_SMEM, st>>>(reinterpret_cast<const uint4 *>(x),
                                                 reinterpret_cast<const uint4 *>(w_img), b1, b2,
                                                 reinterpret_cast<uint4 *>(y), y32, B, T);
  WM_CHECK_LAUNCH("resblock_tc");
  return 0;
}

}  // namespace wm
