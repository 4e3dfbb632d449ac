// 64 -> 64 convolution as an implicit GEMM on the 5th-generation tensor cores (sm_100a):
// tcgen05.mma kind::f16 (bf16 operands, fp32 accumulation in TMEM), operands staged in shared
// memory by bulk asynchronous copies (cp.async.bulk, SASS UBLKCP), persistent warp-specialised CTAs.
//
// Precision: the reference is fp32 and the parity budget (delta 1e-3, per-sample probability
// 1e-3) rules out single-pass bf16 or tf32 operands.  Every operand is
// therefore carried as a bf16 pair  v = hi + lo  and the MMA computes all four partial
// products:  N is doubled to 128 with B = [W_hi | W_lo], and two MMAs per K step feed A_hi
// and A_lo into the same accumulator, so  D[:, 0:64] + D[:, 64:128] = (A_hi + A_lo)(W_hi + W_lo).
//
// Activation layout in HBM ("planar"): per clip 16 planes of (T + 8) rows x 16 bytes:
//   plane c      (c = 0..7)  : bf16 hi of channels 8c..8c+7 of every time step
//   plane 8 + c              : bf16 lo of the same channels
// with 4 zero rows before t = 0 and after t = T-1 (the convolution's zero padding).  A plane is
// exactly the tcgen05 no-swizzle K-major canonical layout (8 rows x 16 B core matrices,
// SBO = 128 B, LBO = plane pitch), so a tile is 16 linear bulk copies, and the operand of tap j
// is the same tile with its start address moved by j rows (16 j bytes).
#include <cuda_bf16.h>

#include "wm_common.h"
#include "wm_tc.cuh"

namespace wm {

using namespace tc;

namespace {

constexpr uint32_t kIdesc = make_idesc(128, 128);
// A_lo multiplies W_hi only (first 64 columns of B): the lo x lo product is below the parity budget
#ifndef WM_FULL_PRODUCTS
constexpr uint32_t kIdescLo = make_idesc(128, 64);
#else
constexpr uint32_t kIdescLo = kIdesc;
#endif
constexpr int PAD = WM_PLANAR_PAD;
constexpr int TILE = 128;

template <int TAPS>
struct Cfg {
  static constexpr int ROWS = TILE + TAPS - 1;
  static constexpr int P = TAPS / 2;
  static constexpr int PLANE_BYTES = ROWS * 16;
  static constexpr int A_STAGE_BYTES = 16 * PLANE_BYTES;
  static constexpr int W_TAP_BYTES = 8 * 128 * 16;
  static constexpr int W_BYTES = TAPS * W_TAP_BYTES;
  static constexpr int NSTAGE = (TAPS == 7) ? 3 : 4;
  static constexpr int BAR_OFF = W_BYTES + NSTAGE * A_STAGE_BYTES;
  static constexpr int SMEM_BYTES = BAR_OFF + 256 + 256;  // barriers + bias
};

}  // namespace

// ---------------------------------------------------------------------------
// The kernel.  Warps 0..15 epilogue (TMEM lane quadrant = warp % 4, 16-channel slice = warp / 4),
// warp 16 producer (bulk copies), warp 17 MMA issuer and TMEM owner.
// ---------------------------------------------------------------------------
template <int TAPS>
__global__ void __launch_bounds__(576, 1)
    conv64_tc_kernel(const uint4 *__restrict__ x, const uint4 *__restrict__ w_img, const float *__restrict__ bias,
                     const uint4 *__restrict__ residual, uint4 *__restrict__ y, float *__restrict__ y32, int B, int T,
                     int relu, const float *__restrict__ res32) {
  using C = Cfg<TAPS>;
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t s_base = smem_u32(smem);
  const uint32_t w_smem = s_base;
  const uint32_t a_smem = s_base + C::W_BYTES;
  const uint32_t bars = s_base + C::BAR_OFF;
  // barrier map (8 bytes each): full[s] 0.., empty[s] NSTAGE.., wbar, tmem_full[2], tmem_empty[2]
  auto full_bar = [&](int s) { return bars + 8 * s; };
  auto empty_bar = [&](int s) { return bars + 8 * (C::NSTAGE + s); };
  const uint32_t wbar = bars + 8 * (2 * C::NSTAGE);
  auto tfull_bar = [&](int a) { return bars + 8 * (2 * C::NSTAGE + 1 + a); };
  auto tempty_bar = [&](int a) { return bars + 8 * (2 * C::NSTAGE + 3 + a); };
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + C::BAR_OFF + 8 * (2 * C::NSTAGE + 5));
  float *bias_s = reinterpret_cast<float *>(smem + C::BAR_OFF + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntile_t = (T + TILE - 1) / TILE;
  const long long ntiles = (long long)B * ntile_t;
  const size_t RP = (size_t)T + 2 * PAD;  // rows per plane

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::NSTAGE; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(wbar, 1);
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 16); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 64) bias_s[threadIdx.x] = bias[threadIdx.x];
  if (warp == 17) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(256)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 16) {
    // ===== producer =====
    if (lane == 0) {
      mbar_arrive_expect_tx(wbar, C::W_BYTES);
      for (int j = 0; j < TAPS; ++j)
        bulk_g2s(w_smem + j * C::W_TAP_BYTES, reinterpret_cast<const uint8_t *>(w_img) + (size_t)j * C::W_TAP_BYTES,
                 C::W_TAP_BYTES, wbar);
      int i = 0;
      TileWalk wk(blockIdx.x, gridDim.x, ntile_t);
      for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++i, wk.next()) {
        const int s = i % C::NSTAGE;
        const uint32_t ph = (i / C::NSTAGE) & 1;
        mbar_wait(empty_bar(s), ph ^ 1);
        mbar_arrive_expect_tx(full_bar(s), C::A_STAGE_BYTES);
        const long long b = wk.b;
        const int t0 = wk.tt * TILE;
        const size_t row0 = (size_t)(t0 + PAD - C::P);
#pragma unroll 4
        for (int p = 0; p < 16; ++p)
          bulk_g2s(a_smem + s * C::A_STAGE_BYTES + p * C::PLANE_BYTES, x + ((size_t)(b * 16 + p) * RP + row0),
                   C::PLANE_BYTES, full_bar(s));
      }
    }
    __syncwarp();
  } else if (warp == 17) {
    // ===== MMA issuer =====  (whole warp walks the pipeline; one elected lane issues)
    const bool issuer = elect_one();
    mbar_wait_warp(wbar, 0);
    const uint64_t b0 = smem_desc(w_smem, 2048, 128);
    int i = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++i) {
      const int s = i % C::NSTAGE;
      const uint32_t ph = (i / C::NSTAGE) & 1;
      const int a = i & 1;
      const uint32_t aph = (i >> 1) & 1;
      mbar_wait_warp(tempty_bar(a), aph ^ 1);
      mbar_wait_warp(full_bar(s), ph);
      tc_fence_after();
      if (issuer) {
        const uint32_t d_tmem = tmem_base + a * 128;
        const uint64_t a0 = smem_desc(a_smem + s * C::A_STAGE_BYTES, C::PLANE_BYTES, 128);
        // descriptors differ only in the 16-byte-granular start address: add compile-time offsets
#pragma unroll
        for (int j = 0; j < TAPS; ++j) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              mma_bf16(d_tmem, a0 + (uint64_t)(((half * 8 + 2 * kk) * C::PLANE_BYTES + j * 16) >> 4),
                       b0 + (uint64_t)((j * C::W_TAP_BYTES + (2 * kk) * 2048) >> 4), half ? kIdescLo : kIdesc,
                       (j | kk | half) != 0 ? 1u : 0u);
            }
          }
        }
        tc_commit(empty_bar(s));   // smem stage is free once these MMAs have read it
        tc_commit(tfull_bar(a));   // accumulator complete
      }
      __syncwarp();
    }
  } else {
    // ===== epilogue =====
    const int q = warp & 3, p = warp >> 2;
    int i = 0;
    TileWalk wk(blockIdx.x, gridDim.x, ntile_t);
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++i, wk.next()) {
      const int a = i & 1;
      const uint32_t aph = (i >> 1) & 1;
      const long long b = wk.b;
      const int t0 = wk.tt * TILE;
      const int t = t0 + q * 32 + lane;
      const bool live = t < T;
      if (y != nullptr && warp == 0 && lane < 2 * PAD) {
        // keep the planes' zero padding rows intact: first / last tile of a clip rewrites them
        const bool head = lane < PAD;
        if (head ? (t0 == 0) : (t0 + TILE >= T)) {
          const size_t zr = head ? (size_t)lane : (size_t)T + lane;
          for (int pl = 0; pl < 16; ++pl) y[((size_t)(b * 16 + pl)) * RP + zr] = make_uint4(0, 0, 0, 0);
        }
      }
      mbar_wait_warp(tfull_bar(a), aph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + a * 128 + ((uint32_t)(q * 32) << 16);
      const size_t prow = (size_t)t + PAD;
      {
        float v1[16], v2[16];
        tmem_ld16(taddr + p * 16, v1);
        tmem_ld16(taddr + 64 + p * 16, v2);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive_warp(tempty_bar(a));   // this thread's slice of the accumulator has been read
        float o[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) o[c] = v1[c] + v2[c] + bias_s[p * 16 + c];
        if (residual != nullptr && live) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int ch = p * 2 + h;
            uint4 rh = __ldg(&residual[((size_t)(b * 16 + ch)) * RP + prow]);
            uint4 rl = __ldg(&residual[((size_t)(b * 16 + 8 + ch)) * RP + prow]);
            float r[8];
            join8(rh, rl, r);
#pragma unroll
            for (int c = 0; c < 8; ++c) o[h * 8 + c] += r[c];
          }
        }
        if (res32 != nullptr && live) {   // fp32 channels-last residual (the training step's gradient accumulation)
          const float4 *src = reinterpret_cast<const float4 *>(res32 + ((size_t)b * T + t) * 64 + p * 16);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float4 r = __ldg(src + c);
            o[4 * c] += r.x; o[4 * c + 1] += r.y; o[4 * c + 2] += r.z; o[4 * c + 3] += r.w;
          }
        }
        if (relu) {
#pragma unroll
          for (int c = 0; c < 16; ++c) o[c] = fmaxf(o[c], 0.0f);
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
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 17) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(256) : "memory");
  }
}

template <int TAPS>
static int launch_conv64_tc_t(const void *x, const void *w_img, const float *bias, const void *residual, void *y,
                              float *y32, int B, int T, int relu, cudaStream_t st, const float *res32) {
  using C = Cfg<TAPS>;
  static bool attr_set = false;
  if (!attr_set) {
    WM_CHECK_CUDA(cudaFuncSetAttribute(conv64_tc_kernel<TAPS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       C::SMEM_BYTES));
    attr_set = true;
  }
  long long ntiles = (long long)B * ((T + TILE - 1) / TILE);
  int grid = (int)(ntiles < sm_count() ? ntiles : sm_count());
  conv64_tc_kernel<TAPS><<<grid, 576, C::SMEM_BYTES, st>>>(
      reinterpret_cast<const uint4 *>(x), reinterpret_cast<const uint4 *>(w_img), bias,
      reinterpret_cast<const uint4 *>(residual), reinterpret_cast<uint4 *>(y), y32, B, T, relu, res32);
  WM_CHECK_LAUNCH("conv64_tc");
  return 0;
}

int launch_conv64_tc(const void *x, const void *w_img, const float *bias, const void *residual, void *y, float *y32,
                     int B, int T, int taps, int relu, cudaStream_t st, const float *res32) {
  if (B == 0 || T == 0) return 0;
  switch (taps) {
    case 3: return launch_conv64_tc_t<3>(x, w_img, bias, residual, y, y32, B, T, relu, st, res32);
    case 7: return launch_conv64_tc_t<7>(x, w_img, bias, residual, y, y32, B, T, relu, st, res32);
    default: set_error("conv64_tc: taps must be 3 or 7 (got %d)", taps); return -1;
  }
}

// ---------------------------------------------------------------------------
// weight image: fp32 w[tap][ci][co] -> bf16 [tap][ci/8][n = 0..127][ci%8], n < 64: hi of co = n,
// n >= 64: lo of co = n - 64   (the B operand [W_hi | W_lo], K-major, no swizzle)
// ---------------------------------------------------------------------------
__global__ void pack_conv64_tc_kernel(const float *__restrict__ w, __nv_bfloat16 *__restrict__ img, int taps) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= taps * 8 * 128 * 8) return;
  int i = e & 7, n = (e >> 3) & 127, c = (e >> 10) & 7, tap = e >> 13;
  float v = w[(tap * 64 + c * 8 + i) * 64 + (n & 63)];
  __nv_bfloat16 hi = __float2bfloat16_rn(v);
  img[e] = n < 64 ? hi : __float2bfloat16_rn(v - __bfloat162float(hi));
}

int launch_pack_conv64_tc(const float *w, void *img, int taps, cudaStream_t st) {
  int n = taps * 8 * 128 * 8;
  pack_conv64_tc_kernel<<<(n + 255) / 256, 256, 0, st>>>(w, reinterpret_cast<__nv_bfloat16 *>(img), taps);
  WM_CHECK_LAUNCH("pack_conv64_tc");
  return 0;
}

// ---------------------------------------------------------------------------
// layout converters and producers of the planar format
// ---------------------------------------------------------------------------
// fp32 channels-last x[b][t][64] (+ per-clip channel vector) -> planar.  One thread per (t, chunk);
// consecutive threads take consecutive t so the 16-byte stores of a warp are contiguous.
__global__ void __launch_bounds__(256)
    to_planar_kernel(const float *__restrict__ x, const float *__restrict__ chan_add, uint4 *__restrict__ y, int T) {
  const int b = blockIdx.z, c = blockIdx.y, t = blockIdx.x * 256 + threadIdx.x;
  if (blockIdx.x == 0 && threadIdx.x < 2 * PAD) {  // the planes' zero padding rows
    const size_t RPz = (size_t)T + 2 * PAD, zr = threadIdx.x < PAD ? threadIdx.x : T + threadIdx.x;
    y[((size_t)b * 16 + c) * RPz + zr] = make_uint4(0, 0, 0, 0);
    y[((size_t)b * 16 + 8 + c) * RPz + zr] = make_uint4(0, 0, 0, 0);
  }
  if (t >= T) return;
  const float4 *src = reinterpret_cast<const float4 *>(x + ((size_t)b * T + t) * 64 + c * 8);
  float4 a0 = __ldg(src), a1 = __ldg(src + 1);
  float v[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
  if (chan_add) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] += chan_add[(size_t)b * 64 + c * 8 + i];
  }
  uint4 hi, lo;
  split8(v, hi, lo);
  const size_t RP = (size_t)T + 2 * PAD;
  y[((size_t)b * 16 + c) * RP + t + PAD] = hi;
  y[((size_t)b * 16 + 8 + c) * RP + t + PAD] = lo;
}

int launch_to_planar(const float *x, const float *chan_add, void *y, int B, int T, cudaStream_t st) {
  if (B == 0 || T == 0) return 0;
  dim3 grid((T + 255) / 256, 8, B);
  to_planar_kernel<<<grid, 256, 0, st>>>(x, chan_add, reinterpret_cast<uint4 *>(y), T);
  WM_CHECK_LAUNCH("to_planar");
  return 0;
}

__global__ void __launch_bounds__(256)
    from_planar_kernel(const uint4 *__restrict__ x, float *__restrict__ y, int T) {
  const int b = blockIdx.z, c = blockIdx.y, t = blockIdx.x * 256 + threadIdx.x;
  if (t >= T) return;
  const size_t RP = (size_t)T + 2 * PAD;
  uint4 hi = __ldg(&x[((size_t)b * 16 + c) * RP + t + PAD]);
  uint4 lo = __ldg(&x[((size_t)b * 16 + 8 + c) * RP + t + PAD]);
  float v[8];
  join8(hi, lo, v);
  float4 *dst = reinterpret_cast<float4 *>(y + ((size_t)b * T + t) * 64 + c * 8);
  dst[0] = make_float4(v[0], v[1], v[2], v[3]);
  dst[1] = make_float4(v[4], v[5], v[6], v[7]);
}

int launch_from_planar(const void *x, float *y, int B, int T, cudaStream_t st) {
  if (B == 0 || T == 0) return 0;
  dim3 grid((T + 255) / 256, 8, B);
  from_planar_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const uint4 *>(x), y, T);
  WM_CHECK_LAUNCH("from_planar");
  return 0;
}

// Conv1d(1,64,7,p=3) writing the planar format directly (py/main16.py:134,177)
__global__ void __launch_bounds__(256)
    conv_in_k7_planar_kernel(const float *__restrict__ s, const float *__restrict__ w, const float *__restrict__ bias,
                             uint4 *__restrict__ y, int T) {
  __shared__ float ss[256 + 8];
  __shared__ __align__(16) float ws[7 * 64 + 64];
  const int b = blockIdx.y, t0 = blockIdx.x * 256, tid = threadIdx.x;
  const float *sb = s + (size_t)b * T;
  for (int i = tid; i < 256 + 6; i += 256) {
    int t = t0 + i - 3;
    ss[i] = (t >= 0 && t < T) ? sb[t] : 0.0f;
  }
  for (int i = tid; i < 7 * 64; i += 256) ws[i] = w[i];
  if (tid < 64) ws[7 * 64 + tid] = bias[tid];
  __syncthreads();
  if (blockIdx.x == 0 && tid < 2 * PAD) {  // the planes' zero padding rows
    const size_t RPz = (size_t)T + 2 * PAD, zr = tid < PAD ? tid : T + tid;
    for (int pl = 0; pl < 16; ++pl) y[((size_t)b * 16 + pl) * RPz + zr] = make_uint4(0, 0, 0, 0);
  }
  const int t = t0 + tid;
  if (t >= T) return;
  float sv[7];
#pragma unroll
  for (int j = 0; j < 7; ++j) sv[j] = ss[tid + j];
  const size_t RP = (size_t)T + 2 * PAD;
#pragma unroll 1
  for (int c = 0; c < 8; ++c) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float a = ws[7 * 64 + c * 8 + i];
#pragma unroll
      for (int j = 0; j < 7; ++j) a = fmaf(sv[j], ws[j * 64 + c * 8 + i], a);
      v[i] = a;
    }
    uint4 hi, lo;
    split8(v, hi, lo);
    y[((size_t)b * 16 + c) * RP + t + PAD] = hi;
    y[((size_t)b * 16 + 8 + c) * RP + t + PAD] = lo;
  }
}

int launch_conv_in_k7_planar(const float *s, const float *w, const float *b, void *y, int B, int T, cudaStream_t st) {
  if (B == 0 || T == 0) return 0;
  dim3 grid((T + 255) / 256, B);
  conv_in_k7_planar_kernel<<<grid, 256, 0, st>>>(s, w, b, reinterpret_cast<uint4 *>(y), T);
  WM_CHECK_LAUNCH("conv_in_k7_planar");
  return 0;
}

}  // namespace wm
