// The FIRST ResBlock of a network fused with the input convolution in front of it:
//     x0 = Conv1d(1,64,7,p=3)(s)                      (py/main16.py:134 / :177)
//     y  = relu( x0 + conv2( relu( conv1(x0) + b1 ) ) + b2 )      (py/main16.py:112-125, BN folded)
// There is no non-linearity between the input convolution and conv1, so their composition is ONE
// 9-tap 1->64 convolution of the waveform,
//     conv1(x0)[t] = B9 + sum_{m<9} W9[m] s[t+m-4],    W9[k+j] += W1[k] . w_in[j],
// (composed on the host in float64, packing.py) except on the first and last sample of a clip, where
// conv1's zero padding of x0 drops one of its three taps; those two rows subtract that tap's composed
// contribution (WK, BK) again.  Both 1->64 convolutions of the waveform are run as ONE small GEMM on the
// tensor pipe: the A operand is the Toeplitz (im2col) tile of the waveform, built on the fly,
//     A[r][m] = s[t0 - 5 + r + m]   (128 rows x K = 16, m >= 9 zero),
// and B = [ W9 | w_in shifted by 2 ] (K = 16 x N = 128), so that accumulator columns 0..63 hold
// conv1(x0) at time t0 - 1 + r (the rows of conv2's operand tile) and columns 64..127 hold x0 at time
// t0 + r (the rows of the OUTPUT tile).  With bf16 hi/lo pairs that is 3 MMAs per tile instead of the
// 24 of a 64-channel conv1, and x0 (4 MB per clip) is never written to or read from HBM.
// The residual never leaves tensor memory either: the conv1 epilogue stores x0 + b_in + b2 into conv2's
// accumulator, and conv2 accumulates on top of it.
//
// Warps 0..15  group 1: D1 (u | x0) -> relu(u + B9) as bf16 hi/lo planes in the intermediate tile U (conv2's
//              A operand); x0 + biases -> D2 (tcgen05.st)
// warps 16..23 group 2: D2 -> ReLU -> planar y; two sub-groups of 4 warps take alternate tiles
// warp 24      producer: Toeplitz tiles of the waveform (bf16 hi/lo), conv2 weights (bulk copy, once)
// warp 25      MMA issue, TMEM owner
// Tile geometry, operand formats and the 3-partial-product bf16 pair scheme are those of
// wm_resblock_tc.cu.
#include <cuda_bf16.h>
#include <string.h>

#include "wm_common.h"
#include "wm_tc.cuh"

namespace wm {

using namespace tc;

namespace {

constexpr int PAD = WM_PLANAR_PAD;
constexpr int TO = 126;                    // output rows per tile
constexpr int XROWS = 130;                 // rows of an intermediate tile buffer (128 used + tap reach)
constexpr int PLANE_B = XROWS * 16;
constexpr int TILE_B = 16 * PLANE_B;       // 33 280
constexpr int W_TAP_B = 8 * 128 * 16;
constexpr int W_IMG_B = 3 * W_TAP_B;       // 49 152
constexpr int NU = 2;                      // intermediate tiles in flight
constexpr int AT_B = 2 * 128 * 16;         // one Toeplitz operand (hi or lo): [k chunk 2][row 128][8 x bf16]
constexpr int BT_B = 2 * 128 * 16;         // one B operand (hi or lo):        [k chunk 2][col 128][8 x bf16]
constexpr int OFF_W = 0;
constexpr int OFF_BT = OFF_W + W_IMG_B;            // B_hi, B_lo
constexpr int OFF_AT = OFF_BT + 2 * BT_B;          // 2 stages x (A_hi, A_lo)
constexpr int OFF_U = OFF_AT + 2 * 2 * AT_B;
constexpr int OFF_BIAS = OFF_U + NU * TILE_B;      // b9[64], (b_in + b2)[64]
constexpr int OFF_WIN = OFF_BIAS + 512;            // the producer's waveform window (160 floats)
constexpr int OFF_BAR = OFF_WIN + 640;
constexpr int RBI_SMEM = OFF_BAR + 160;
static_assert(RBI_SMEM <= 232448, "shared memory budget");
constexpr uint32_t kIdesc = make_idesc(128, 128);
constexpr uint32_t kIdescLo = make_idesc(128, 64);     // A_lo x W_hi only
constexpr int N_G1 = 512;                  // group 1: 16 warps = 4 TMEM lane quadrants x 4 slices of 16 channels
constexpr int N_G2 = 256;                  // group 2: 2 sub-groups of 4 warps
constexpr int W_PROD = (N_G1 + N_G2) / 32, W_MMA = W_PROD + 1, RBI_THREADS = N_G1 + N_G2 + 64;

}  // namespace

// w9g: device w9[9][64], b9[64]; wing: device w_in[7][64], b_in[64]; fin: device WM_FIN_* block (edge rows)
template <bool PROF>
__global__ void __launch_bounds__(RBI_THREADS, 1)
    resblock_in_tc_kernel(const float *__restrict__ s, const float *__restrict__ w9g, const float *__restrict__ wing,
                          const float *__restrict__ fin, const uint4 *__restrict__ w_img2, const float *__restrict__ b2,
                          uint4 *__restrict__ y, int B, int T, long long *__restrict__ prof) {
  // PROF: per-phase cycle sums of block 0 (tools/resblock_in_profile.py); the production instantiation has none
  auto clk = [&]() -> long long { return PROF ? clock64() : 0; };
  const bool pf0 = PROF && prof != nullptr && blockIdx.x == 0;
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t s_base = smem_u32(smem);
  const uint32_t w_smem = s_base + OFF_W, bt_smem = s_base + OFF_BT, at_smem = s_base + OFF_AT, u_smem = s_base + OFF_U;
  const uint32_t bars = s_base + OFF_BAR;
  enum { WBAR = 0, A_FULL = 1, A_EMPTY = 3, D1_FULL = 5, U_FULL = 7, U_EMPTY = 9, D2_FULL = 11, D2_EMPTY = 13, NBAR = 15 };
  auto bar = [&](int i) { return bars + 8 * i; };
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + OFF_BAR + 8 * NBAR);
  float *bias_s = reinterpret_cast<float *>(smem + OFF_BIAS);   // [0..63] b9, [64..127] b_in + b2

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntile_t = (T + TO - 1) / TO;
  const long long ntiles = (long long)B * ntile_t;
  const long long my_tiles = ntiles > blockIdx.x ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const size_t RP = (size_t)T + 2 * PAD;

  if (threadIdx.x == 0) {
    mbar_init(bar(WBAR), 1);
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar(A_FULL + a), 1); mbar_init(bar(A_EMPTY + a), 1); mbar_init(bar(D1_FULL + a), 1);
      mbar_init(bar(U_FULL + a), N_G1 / 32); mbar_init(bar(U_EMPTY + a), 1);
      mbar_init(bar(D2_FULL + a), 1); mbar_init(bar(D2_EMPTY + a), 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 128)
    bias_s[threadIdx.x] = threadIdx.x < 64 ? w9g[9 * 64 + threadIdx.x]
                                           : wing[7 * 64 + threadIdx.x - 64] + b2[threadIdx.x - 64];
  // B operand of the Toeplitz GEMM, bf16 hi and lo: element (k, n), k < 16, n < 128:
  //   n < 64: W9[k][n] (k < 9);   n >= 64: w_in[k - 2][n - 64] (2 <= k < 9);   else 0
  for (int e = threadIdx.x; e < 16 * 128; e += RBI_THREADS) {
    const int k = e >> 7, n = e & 127;
    float v = 0.0f;
    if (n < 64) { if (k < 9) v = w9g[k * 64 + n]; }
    else if (k >= 2 && k < 9) v = wing[(k - 2) * 64 + n - 64];
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
    const int off = (k >> 3) * 2048 + n * 16 + (k & 7) * 2;
    *reinterpret_cast<__nv_bfloat16 *>(smem + OFF_BT + off) = hi;
    *reinterpret_cast<__nv_bfloat16 *>(smem + OFF_BT + BT_B + off) = lo;
  }
  if (warp == W_MMA) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == W_PROD) {
    // ===== producer: conv2 weights once; per tile the Toeplitz tile of the waveform (4 rows per lane) =====
    if (lane == 0) {
      mbar_arrive_expect_tx(bar(WBAR), W_IMG_B);
      for (int j = 0; j < 3; ++j)
        bulk_g2s(w_smem + j * W_TAP_B, reinterpret_cast<const uint8_t *>(w_img2) + (size_t)j * W_TAP_B, W_TAP_B, bar(WBAR));
    }
    // The tile's 136 waveform samples s[t0-5 .. t0+130] are fetched ONCE (coalesced, one tile ahead, so the HBM
    // latency hides behind the previous build), parked in a private shared-memory window, and the 128 Toeplitz
    // rows are cut from that window (measured: 36 dependent global loads per lane made this warp pace the kernel).
    float *swin = reinterpret_cast<float *>(smem + OFF_WIN);
    float wreg[5];
    TileWalk wk(blockIdx.x, gridDim.x, ntile_t);
    auto fetch_window = [&]() {   // the window of the walk's current tile; advances the walk
      const float *sb = s + (size_t)wk.b * T;
      const int ts0 = wk.tt * TO - 5;
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        const int ts = ts0 + lane + 32 * k;
        wreg[k] = (ts >= 0 && ts < T) ? __ldg(sb + ts) : 0.0f;
      }
      wk.next();
    };
    if (my_tiles > 0) fetch_window();
    for (long long i = 0; i < my_tiles; ++i) {
      const int st = (int)(i & 1);
#pragma unroll
      for (int k = 0; k < 5; ++k) swin[lane + 32 * k] = wreg[k];
      __syncwarp();
      if (i + 1 < my_tiles) fetch_window();
      const long long p0 = clk();
      if (i >= 2) mbar_wait_warp(bar(A_EMPTY + st), (uint32_t)(((i >> 1) - 1) & 1));
      const long long p1 = clk();
      uint8_t *ah = smem + OFF_AT + st * 2 * AT_B, *al = ah + AT_B;
#pragma unroll
      for (int rr = 0; rr < 4; ++rr) {
        const int r = rr * 32 + lane;
        float v[10];
#pragma unroll
        for (int m = 0; m < 9; ++m) v[m] = swin[r + m];
        v[9] = 0.0f;
        uint32_t h[5], l[5];
#pragma unroll
        for (int k = 0; k < 5; ++k) split2(v[2 * k], v[2 * k + 1], h[k], l[k]);
        *reinterpret_cast<uint4 *>(ah + r * 16) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4 *>(ah + 2048 + r * 16) = make_uint4(h[4], 0, 0, 0);
        *reinterpret_cast<uint4 *>(al + r * 16) = make_uint4(l[0], l[1], l[2], l[3]);
        *reinterpret_cast<uint4 *>(al + 2048 + r * 16) = make_uint4(l[4], 0, 0, 0);
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(A_FULL + st));
      if (pf0 && lane == 0) { prof[0] += p1 - p0; prof[1] += clk() - p1; }
    }
  } else if (warp == W_MMA) {
    // ===== MMA issuer =====
    if (elect_one()) {
      const uint64_t bt_hi = smem_desc(bt_smem, 2048, 128), bt_lo = smem_desc(bt_smem + BT_B, 2048, 128);
      auto gemm1 = [&](int st, uint32_t d_tmem) {   // D1 = A_hi B_hi + A_lo B_hi + A_hi B_lo   (K = 16 each)
        const uint64_t a_hi = smem_desc(at_smem + st * 2 * AT_B, 2048, 128);
        const uint64_t a_lo = smem_desc(at_smem + st * 2 * AT_B + AT_B, 2048, 128);
        mma_bf16(d_tmem, a_hi, bt_hi, kIdesc, 0u);
        mma_bf16(d_tmem, a_lo, bt_hi, kIdesc, 1u);
        mma_bf16(d_tmem, a_hi, bt_lo, kIdesc, 1u);
      };
      auto conv2 = [&](uint32_t a_tile, uint32_t d_tmem) {   // accumulates on the residual stored by group 1
        const uint64_t a0 = smem_desc(a_tile, PLANE_B, 128), b0 = smem_desc(w_smem, 2048, 128);
#pragma unroll
        for (int j = 0; j < 3; ++j) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              mma_bf16(d_tmem, a0 + (uint64_t)(((half * 8 + 2 * kk) * PLANE_B + j * 16) >> 4),
                       b0 + (uint64_t)((j * W_TAP_B + (2 * kk) * 2048) >> 4), half ? kIdescLo : kIdesc, 1u);
            }
          }
        }
      };
      mbar_wait(bar(WBAR), 0);
      if (my_tiles > 0) {
        mbar_wait(bar(A_FULL), 0);
        tc_fence_after();
        gemm1(0, tmem);
        tc_commit(bar(D1_FULL));
        tc_commit(bar(A_EMPTY));
      }
      for (long long i = 0; i < my_tiles; ++i) {
        const int a = (int)(i & 1);
        long long m0 = clk(), m1 = m0, m2 = m0;
        if (i + 1 < my_tiles) {   // Toeplitz GEMM of the next tile; D1[an] was drained by group 1 of tile i-1 (u_full(i-1))
          const long long n = i + 1;
          const int an = (int)(n & 1);
          mbar_wait(bar(A_FULL + an), (uint32_t)((n >> 1) & 1));
          m1 = clk();
          tc_fence_after();
          gemm1(an, tmem + an * 128);
          tc_commit(bar(D1_FULL + an));
          tc_commit(bar(A_EMPTY + an));
          m2 = clk();
        }
        mbar_wait(bar(U_FULL + a), (uint32_t)((i >> 1) & 1));    // U[a] written and D2[a] initialised with the residual
        const long long m3 = clk();
        tc_fence_after();
        conv2(u_smem + a * TILE_B, tmem + 256 + a * 128);
        tc_commit(bar(D2_FULL + a));
        if (pf0) { prof[2] += m1 - m0; prof[3] += m2 - m1; prof[4] += m3 - m2; prof[5] += clk() - m3; prof[15] = my_tiles; }
      }
    }
    __syncwarp();
  } else if (warp < N_G1 / 32) {
    // ===== group 1: D1 -> U tile (relu(conv1 + B9), bf16 hi/lo) and residual -> D2 =====
    // (this stage, not the tensor pipe, paces the kernel: 16 warps, 16 channels of one row per thread)
    const int q = warp & 3, p = warp >> 2;             // TMEM lane quadrant, 16-channel slice
    const int row = q * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    TileWalk wk(blockIdx.x, gridDim.x, ntile_t);
    for (long long i = 0; i < my_tiles; ++i, wk.next()) {
      const long long b = wk.b;
      const int t0 = wk.tt * TO;
      const int a = (int)(i & 1);
      const int tu = t0 - 1 + row;
      const bool inside = tu >= 0 && tu < T;     // conv2 zero-pads the intermediate feature map
      const bool edge = inside && (tu == 0 || tu == T - 1);
      uint8_t *us = smem + OFF_U + a * TILE_B + row * 16;
      const long long g0 = clk();
      mbar_wait_warp(bar(D1_FULL + a), (uint32_t)((i >> 1) & 1));
      const long long g1c = clk();
      // (no separate wait for "conv2(i-2) has finished reading U[a]": the D2_EMPTY wait below, made before the first
      // store into U[a], implies it — tile i-2's epilogue drains D2[a] only after those MMAs completed)
      const long long g2c = clk();
      long long g3c = g2c, g4c = g2c;
      tc_fence_after();
      const uint32_t t1 = tmem + a * 128 + lane_off, t2 = tmem + 256 + a * 128 + lane_off;
      {
        float u[16], xr[16];
        tmem_ld16(t1 + p * 16, u);
        tmem_ld16(t1 + 64 + p * 16, xr);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 16; ++c) { u[c] += bias_s[p * 16 + c]; xr[c] += bias_s[64 + p * 16 + c]; }
        // residual x0 + b_in + b2 -> conv2's accumulator (columns of the hi product; the lo-product columns start at 0).
        // Issued BEFORE the u part so that the tensor-memory store's latency (tcgen05.wait::st below took ~900 cycles
        // on this group's critical path) overlaps the ReLU / split / shared-memory stores of the intermediate tile.
        g3c = clk();
        if (i >= 2) {
          mbar_wait_warp(bar(D2_EMPTY + a), (uint32_t)(((i >> 1) - 1) & 1));   // tile i-2's epilogue has drained D2[a]
          tc_fence_after();
        }
        g4c = clk();
        tmem_st16(t2 + p * 16, xr);
#pragma unroll
        for (int c = 0; c < 16; ++c) xr[c] = 0.0f;
        tmem_st16(t2 + 64 + p * 16, xr);
        if (edge) {   // first / last sample of the clip: take back conv1's tap that fell on the zero padding of x0
          const float *sb = s + (size_t)b * T;
          const bool drop0 = tu == 0, drop2 = tu == T - 1;
#pragma unroll
          for (int c = 0; c < 16; ++c) {      // unrolled: u[] must stay in registers
            const int co = p * 16 + c;
            float d = (drop0 ? fin[WM_FIN_BK + co] : 0.0f) + (drop2 ? fin[WM_FIN_BK + 2 * 64 + co] : 0.0f);
#pragma unroll 1
            for (int j = 0; j < 7; ++j) {
              const int ts0 = tu + j - 4, ts2 = tu + j - 2;     // sample under composed tap (k = 0, j) / (k = 2, j)
              if (drop0 && ts0 >= 0 && ts0 < T) d = fmaf(__ldg(sb + ts0), fin[WM_FIN_WK + j * 64 + co], d);
              if (drop2 && ts2 >= 0 && ts2 < T) d = fmaf(__ldg(sb + ts2), fin[WM_FIN_WK + (14 + j) * 64 + co], d);
            }
            u[c] -= d;
          }
        }
#pragma unroll
        for (int c = 0; c < 16; ++c) u[c] = inside ? fmaxf(u[c], 0.0f) : 0.0f;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int ch = p * 2 + h;
          uint4 hi, lo;
          split8(u + h * 8, hi, lo);
          *reinterpret_cast<uint4 *>(us + ch * PLANE_B) = hi;
          *reinterpret_cast<uint4 *>(us + (8 + ch) * PLANE_B) = lo;
        }
      }
      tmem_st_wait();
      tc_fence_before();
      fence_async_smem();
      mbar_arrive_warp(bar(U_FULL + a));
      if (pf0 && threadIdx.x == 0) {
        prof[6] += g1c - g0; prof[7] += g2c - g1c; prof[8] += g3c - g2c; prof[9] += g4c - g3c; prof[10] += clk() - g4c;
      }
    }
  } else {
    // ===== group 2: D2 (residual + conv2, hi and lo products) -> ReLU -> y =====
    const int w2 = warp - N_G1 / 32;
    const int q = w2 & 3, g = w2 >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    TileWalk wk(blockIdx.x + (long long)g * gridDim.x, 2LL * gridDim.x, ntile_t);
    for (long long i = g; i < my_tiles; i += 2, wk.next()) {
      const long long b = wk.b;
      const int t0 = wk.tt * TO;
      const int t = t0 + row;
      const bool live = row < TO && t < T;
      const size_t prow = (size_t)t + PAD;
      if (q == 0 && lane < 2 * PAD) {  // the planes' zero padding rows
        const bool head = lane < PAD;
        if (head ? (t0 == 0) : (t0 + TO >= T)) {
          const size_t zr = head ? (size_t)lane : (size_t)T + lane;
          for (int pl = 0; pl < 16; ++pl) y[((size_t)(b * 16 + pl)) * RP + zr] = make_uint4(0, 0, 0, 0);
        }
      }
      const long long h0 = clk();
      mbar_wait_warp(bar(D2_FULL + g), (uint32_t)((i >> 1) & 1));
      const long long h1 = clk();
      long long h2 = h1;
      tc_fence_after();
      const uint32_t taddr = tmem + 256 + g * 128 + lane_off;
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        float v1[16], v2[16], o[16];
        tmem_ld16(taddr + p * 16, v1);
        tmem_ld16(taddr + 64 + p * 16, v2);
        tmem_ld_wait();
        if (p == 3) {
          tc_fence_before();
          mbar_arrive_warp(bar(D2_EMPTY + g));     // D2[g] may be re-initialised for tile i+2
          h2 = clk();
        }
#pragma unroll
        for (int c = 0; c < 16; ++c) o[c] = fmaxf(v1[c] + v2[c], 0.0f);
        if (live) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int ch = p * 2 + h;
            uint4 hi, lo;
            split8(o + h * 8, hi, lo);
            y[((size_t)(b * 16 + ch)) * RP + prow] = hi;
            y[((size_t)(b * 16 + 8 + ch)) * RP + prow] = lo;
          }
        }
      }
      if (pf0 && w2 == 0 && lane == 0) { prof[11] += h1 - h0; prof[12] += h2 - h1; prof[13] += clk() - h2; }
    }
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
  }
}

// w9b: device block w9[9][64] then b9[64] (WM_FIN_W9 of the blob); winb: device w_in[7][64] then b_in[64];
// fin: device WM_FIN_* block; w_img2 / b2: conv2 image and bias of the ResBlock; s[B][T] -> y planar
int launch_resblock_in_tc(const float *s, const float *w9b, const float *winb, const float *fin, const void *w_img2,
                          const float *b2, void *y, int B, int T, cudaStream_t st) {
  if (B == 0 || T == 0) return 0;
  static bool attr_set = false;
  if (!attr_set) {
    WM_CHECK_CUDA(cudaFuncSetAttribute(resblock_in_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, RBI_SMEM));
    WM_CHECK_CUDA(cudaFuncSetAttribute(resblock_in_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, RBI_SMEM));
    attr_set = true;
  }
  long long ntiles = (long long)B * ((T + TO - 1) / TO);
  int grid = (int)(ntiles < sm_count() ? ntiles : sm_count());
  if (get_profile_buffer() != nullptr)   // developer build with per-phase cycle counters
    resblock_in_tc_kernel<true><<<grid, RBI_THREADS, RBI_SMEM, st>>>(
        s, w9b, winb, fin, reinterpret_cast<const uint4 *>(w_img2), b2, reinterpret_cast<uint4 *>(y), B, T,
        get_profile_buffer());
  else
    resblock_in_tc_kernel<false><<<grid, RBI_THREADS, RBI_SMEM, st>>>(
        s, w9b, winb, fin, reinterpret_cast<const uint4 *>(w_img2), b2, reinterpret_cast<uint4 *>(y), B, T, nullptr);
  WM_CHECK_LAUNCH("resblock_in_tc");
  return 0;
}

}  // namespace wm
