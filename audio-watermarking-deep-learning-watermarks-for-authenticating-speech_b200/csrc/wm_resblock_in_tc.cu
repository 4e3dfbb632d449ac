// The FIRST ResBlock of a network fused with the input convolution in front of it:
//     x0 = Conv1d(1,64,7,p=3)(s)                      (py/main16.py:134 / :177)
//     y  = relu( x0 + conv2( relu( conv1(x0) + b1 ) ) + b2 )      (py/main16.py:112-125, BN folded)
// There is no non-linearity between the input convolution and conv1, so their composition is ONE
// 9-tap 1->64 convolution of the waveform,
//     conv1(x0)[t] = B9 + sum_{m<9} W9[m] s[t+m-4],    W9[k+j] += W1[k] . w_in[j],
// (composed on the host in float64, packing.py) except on the first and last sample of a clip, where
// conv1's zero padding of x0 drops one of its three taps; those two rows subtract that tap's composed
// contribution (WK, BK) again.  So this kernel never materialises x0 in HBM (4 MB per clip written and
// read back), runs conv1 on the CUDA cores in fp32 straight from the waveform and only conv2 on the
// tensor pipe: half the MMAs and half the shared-memory operand traffic of the general ResBlock kernel,
// which is what that kernel is bound by.
//
// Warps 0..7   group 1 (producers): warp w owns rows 16w..16w+15 of the tile, lane = channel pair, so the
//              9 + 7 taps of its two channels stay in REGISTERS (packed f32x2) for the whole kernel; per
//              row 16 FFMA2 give u = conv1(conv_in(s)) and x0 = conv_in(s); relu(u) goes to the
//              intermediate tile U (bf16 hi/lo planes, conv2's A operand), x0 (fp32) to a side tile
// warps 8..15  group 2: conv2 accumulator + b2 + x0 (from the side tile) -> ReLU -> planar y; two
//              sub-groups of 4 warps take alternate tiles
// warp 16      conv2 weights (bulk copy, once), MMA issue, TMEM owner
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
constexpr int NX = 3;                      // residual side tiles in flight
constexpr int X0_PITCH = 272;              // bytes per row of the side tile: 64 fp32 + 16 (bank spread for LDS.128)
constexpr int X0_B = 128 * X0_PITCH;       // 34 816
constexpr int OFF_W = 0;
constexpr int OFF_U = OFF_W + W_IMG_B;
constexpr int OFF_X0 = OFF_U + NU * TILE_B;
constexpr int OFF_BIAS = OFF_X0 + NX * X0_B;
constexpr int OFF_BAR = OFF_BIAS + 256;
constexpr int RBI_SMEM = OFF_BAR + 128;
static_assert(RBI_SMEM <= 232448, "shared memory budget");
constexpr uint32_t kIdesc = make_idesc(128, 128);
constexpr uint32_t kIdescLo = make_idesc(128, 64);     // A_lo x W_hi only
constexpr int N_GRP = 256;
constexpr int W_MMA = 2 * N_GRP / 32, RBI_THREADS = 2 * N_GRP + 32;

}  // namespace

// in: device block  w9[9][64], b9[64], w_in[7][64], b_in[64]  (WM_FIN_W9.. and the input convolution)
// fin: device pointer to the WM_FIN_* block of the blob (boundary rows read WK / BK from it)
__global__ void __launch_bounds__(RBI_THREADS, 1)
    resblock_in_tc_kernel(const float *__restrict__ s, const float *__restrict__ w9g, const float *__restrict__ wing,
                          const float *__restrict__ fin, const uint4 *__restrict__ w_img2, const float *__restrict__ b2,
                          uint4 *__restrict__ y, int B, int T) {
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t s_base = smem_u32(smem);
  const uint32_t w_smem = s_base + OFF_W, u_smem = s_base + OFF_U;
  const uint32_t bars = s_base + OFF_BAR;
  enum { WBAR = 0, U_FULL = 1, U_EMPTY = U_FULL + NU, D2_FULL = U_EMPTY + NU, D2_EMPTY = D2_FULL + 2,
         X_EMPTY = D2_EMPTY + 2, NBAR = X_EMPTY + NX };
  static_assert(8 * NBAR + 8 <= 128, "barrier block");
  auto bar = [&](int i) { return bars + 8 * i; };
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + OFF_BAR + 8 * NBAR);
  float *bias_s = reinterpret_cast<float *>(smem + OFF_BIAS);   // b2

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntile_t = (T + TO - 1) / TO;
  const long long ntiles = (long long)B * ntile_t;
  const long long my_tiles = ntiles > blockIdx.x ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const size_t RP = (size_t)T + 2 * PAD;

  if (threadIdx.x == 0) {
    mbar_init(bar(WBAR), 1);
    for (int a = 0; a < NU; ++a) { mbar_init(bar(U_FULL + a), N_GRP / 32); mbar_init(bar(U_EMPTY + a), 1); }
    for (int g = 0; g < 2; ++g) { mbar_init(bar(D2_FULL + g), 1); mbar_init(bar(D2_EMPTY + g), 4); }
    for (int x = 0; x < NX; ++x) mbar_init(bar(X_EMPTY + x), 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 64) bias_s[threadIdx.x] = b2[threadIdx.x];
  if (warp == W_MMA) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(256)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == W_MMA) {
    // ===== conv2: weights once, then 24 MMAs per tile =====
    if (elect_one()) {
      mbar_arrive_expect_tx(bar(WBAR), W_IMG_B);
      for (int j = 0; j < 3; ++j)
        bulk_g2s(w_smem + j * W_TAP_B, reinterpret_cast<const uint8_t *>(w_img2) + (size_t)j * W_TAP_B, W_TAP_B, bar(WBAR));
      mbar_wait(bar(WBAR), 0);
      const uint64_t b0 = smem_desc(w_smem, 2048, 128);
      for (long long i = 0; i < my_tiles; ++i) {
        const int a = (int)(i % NU), g = (int)(i & 1);
        mbar_wait(bar(U_FULL + a), (uint32_t)((i / NU) & 1));
        if (i >= 2) mbar_wait(bar(D2_EMPTY + g), (uint32_t)(((i >> 1) - 1) & 1));
        tc_fence_after();
        const uint64_t a0 = smem_desc(u_smem + a * TILE_B, PLANE_B, 128);
        const uint32_t d_tmem = tmem + g * 128;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              mma_bf16(d_tmem, a0 + (uint64_t)(((half * 8 + 2 * kk) * PLANE_B + j * 16) >> 4),
                       b0 + (uint64_t)((j * W_TAP_B + (2 * kk) * 2048) >> 4), half ? kIdescLo : kIdesc,
                       (j | kk | half) != 0 ? 1u : 0u);
            }
          }
        }
        tc_commit(bar(D2_FULL + g));
        tc_commit(bar(U_EMPTY + a));
      }
    }
    __syncwarp();
  } else if (warp < N_GRP / 32) {
    // ===== group 1: rows 16*warp .. +15 of the tile (time t0 - 1 + row), channels 2*lane, 2*lane + 1 =====
    const int c0 = 2 * lane;
    f32x2 w9[9], wi[7];
#pragma unroll
    for (int m = 0; m < 9; ++m) w9[m] = pk2(__ldg(w9g + m * 64 + c0), __ldg(w9g + m * 64 + c0 + 1));
#pragma unroll
    for (int j = 0; j < 7; ++j) wi[j] = pk2(__ldg(wing + j * 64 + c0), __ldg(wing + j * 64 + c0 + 1));
    const f32x2 b9 = pk2(__ldg(w9g + 9 * 64 + c0), __ldg(w9g + 9 * 64 + c0 + 1));
    const f32x2 bi = pk2(__ldg(wing + 7 * 64 + c0), __ldg(wing + 7 * 64 + c0 + 1));
    // this lane's 4 bytes inside the 16-byte row chunk of plane lane / 4 (hi) and 8 + lane / 4 (lo)
    const int u_off = (lane >> 2) * PLANE_B + (lane & 3) * 4;
    for (long long i = 0; i < my_tiles; ++i) {
      const long long tile = blockIdx.x + i * gridDim.x;
      const long long b = tile / ntile_t;
      const int t0 = (int)(tile % ntile_t) * TO;
      const int a = (int)(i % NU), xb = (int)(i % NX);
      const int tu0 = t0 - 1 + 16 * warp;                 // time of this warp's first row
      const float *sb = s + (size_t)b * T;
      float sw[24];                                        // s[tu0 - 4 .. tu0 + 19], zero outside the clip
#pragma unroll
      for (int k = 0; k < 24; ++k) {
        const int ts = tu0 - 4 + k;
        sw[k] = (ts >= 0 && ts < T) ? __ldg(sb + ts) : 0.0f;
      }
      if (i >= NU) mbar_wait_warp(bar(U_EMPTY + a), (uint32_t)(((i / NU) - 1) & 1));   // conv2(i-NU) has read U[a]
      if (i >= NX) mbar_wait_warp(bar(X_EMPTY + xb), (uint32_t)(((i / NX) - 1) & 1));  // epilogue(i-NX) has read X0[xb]
      uint8_t *us = smem + OFF_U + a * TILE_B + (16 * warp) * 16 + u_off;
      uint8_t *xs = smem + OFF_X0 + xb * X0_B + (16 * warp) * X0_PITCH + c0 * 4;
#pragma unroll
      for (int rr = 0; rr < 16; ++rr) {
        const int tu = tu0 + rr;
        f32x2 sp[9];
#pragma unroll
        for (int m = 0; m < 9; ++m) sp[m] = pk2(sw[rr + m], sw[rr + m]);
        f32x2 au = b9, ax = bi;
#pragma unroll
        for (int m = 0; m < 9; ++m) au = fma2(sp[m], w9[m], au);
#pragma unroll
        for (int j = 0; j < 7; ++j) ax = fma2(sp[j + 1], wi[j], ax);
        float u0, u1;
        upk2(au, u0, u1);
        const bool inside = tu >= 0 && tu < T;             // conv2 zero-pads the intermediate feature map
        u0 = inside ? fmaxf(u0, 0.0f) : 0.0f;
        u1 = inside ? fmaxf(u1, 0.0f) : 0.0f;
        uint32_t hi, lo;
        split2(u0, u1, hi, lo);
        *reinterpret_cast<uint32_t *>(us + rr * 16) = hi;
        *reinterpret_cast<uint32_t *>(us + rr * 16 + 8 * PLANE_B) = lo;
        *reinterpret_cast<f32x2 *>(xs + rr * X0_PITCH) = ax;
      }
      // First and last sample of the clip (at most two rows per clip, warp-uniform test): conv1's tap on the zero
      // padding of x0 is not part of the result; redo that row without it.  Kept out of the row loop so that the
      // 16 rows' FMA chains interleave freely.
#pragma unroll 1
      for (int e = 0; e < 2; ++e) {
        const int te = e == 0 ? 0 : T - 1;
        const int rr = te - tu0;
        if (rr < 0 || rr >= 16 || (e == 1 && T == 1)) continue;
        const bool drop0 = te == 0, drop2 = te == T - 1;
        float u0 = __ldg(w9g + 9 * 64 + c0), u1 = __ldg(w9g + 9 * 64 + c0 + 1);
        for (int m = 0; m < 9; ++m) {
          const int ts = te + m - 4;
          const float sv = (ts >= 0 && ts < T) ? __ldg(sb + ts) : 0.0f;
          u0 = fmaf(sv, __ldg(w9g + m * 64 + c0), u0);
          u1 = fmaf(sv, __ldg(w9g + m * 64 + c0 + 1), u1);
          for (int k = 0; k < 3; k += 2) {          // conv1 taps 0 and 2: composed tap index m = k + j
            const int j = m - k;
            if (j < 0 || j > 6 || !(k == 0 ? drop0 : drop2)) continue;
            u0 -= sv * fin[WM_FIN_WK + (k * 7 + j) * 64 + c0];
            u1 -= sv * fin[WM_FIN_WK + (k * 7 + j) * 64 + c0 + 1];
          }
        }
        if (drop0) { u0 -= fin[WM_FIN_BK + c0]; u1 -= fin[WM_FIN_BK + c0 + 1]; }
        if (drop2) { u0 -= fin[WM_FIN_BK + 2 * 64 + c0]; u1 -= fin[WM_FIN_BK + 2 * 64 + c0 + 1]; }
        uint32_t hi, lo;
        split2(fmaxf(u0, 0.0f), fmaxf(u1, 0.0f), hi, lo);
        *reinterpret_cast<uint32_t *>(us + rr * 16) = hi;
        *reinterpret_cast<uint32_t *>(us + rr * 16 + 8 * PLANE_B) = lo;
      }
      fence_async_smem();
      mbar_arrive_warp(bar(U_FULL + a));
    }
  } else {
    // ===== group 2: conv2 accumulator + b2 + x0 -> ReLU -> y =====
    const int w2 = warp - N_GRP / 32;
    const int q = w2 & 3, g = w2 >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    for (long long i = g; i < my_tiles; i += 2) {
      const long long tile = blockIdx.x + i * gridDim.x;
      const long long b = tile / ntile_t;
      const int t0 = (int)(tile % ntile_t) * TO;
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
      // x0 of output row `row` is row + 1 of the side tile (the tile starts one sample early)
      const float4 *xr = reinterpret_cast<const float4 *>(smem + OFF_X0 + (int)(i % NX) * X0_B +
                                                          (row < 127 ? row + 1 : 127) * X0_PITCH);
      mbar_wait_warp(bar(D2_FULL + g), (uint32_t)((i >> 1) & 1));
      tc_fence_after();
      const uint32_t taddr = tmem + g * 128 + lane_off;
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        float v1[16], v2[16], o[16];
        tmem_ld16(taddr + p * 16, v1);
        tmem_ld16(taddr + 64 + p * 16, v2);
        float4 xv[4];
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) xv[c4] = xr[p * 4 + c4];
        tmem_ld_wait();
        if (p == 3) {
          tc_fence_before();
          mbar_arrive_warp(bar(D2_EMPTY + g));              // D2[g] may be overwritten by conv2(i+2)
          mbar_arrive_warp(bar(X_EMPTY + (int)(i % NX)));   // and the side tile by the producers of tile i+NX
        }
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          const float r4[4] = {xv[c4].x, xv[c4].y, xv[c4].z, xv[c4].w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int c = c4 * 4 + k;
            o[c] = fmaxf(v1[c] + v2[c] + bias_s[p * 16 + c] + r4[k], 0.0f);
          }
        }
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
    }
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(256) : "memory");
  }
}

// w9b: device block w9[9][64] then b9[64] (WM_FIN_W9 of the blob); winb: device w_in[7][64] then b_in[64];
// fin: device WM_FIN_* block; w_img2 / b2: conv2 image and bias of the ResBlock; s[B][T] -> y planar
int launch_resblock_in_tc(const float *s, const float *w9b, const float *winb, const float *fin, const void *w_img2,
                          const float *b2, void *y, int B, int T, cudaStream_t st) {
  if (B == 0 || T == 0) return 0;
  static bool attr_set = false;
  if (!attr_set) {
    WM_CHECK_CUDA(cudaFuncSetAttribute(resblock_in_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RBI_SMEM));
    attr_set = true;
  }
  long long ntiles = (long long)B * ((T + TO - 1) / TO);
  int grid = (int)(ntiles < sm_count() ? ntiles : sm_count());
  resblock_in_tc_kernel<<<grid, RBI_THREADS, RBI_SMEM, st>>>(s, w9b, winb, fin, reinterpret_cast<const uint4 *>(w_img2), b2,
                                                            reinterpret_cast<uint4 *>(y), B, T);
  WM_CHECK_LAUNCH("resblock_in_tc");
  return 0;
}

}  // namespace wm
