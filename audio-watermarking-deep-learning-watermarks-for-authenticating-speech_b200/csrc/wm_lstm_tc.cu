// nn.LSTM(64,64,batch_first=True), zero initial state (py/main16.py:138,153) as a persistent
// tensor-core kernel (sm_100a).  One CTA owns 32 clips for all T steps.
//
// Per step the gate pre-activations are computed TRANSPOSED,
//     G^T[256 gate rows, clips] = W_ih . x_t^T + W_hh . h_{t-1}^T ,
// so the (constant) weights are the A operand and stay RESIDENT IN TENSOR MEMORY for the whole
// kernel (tcgen05.mma with A in TMEM), while x_t / h_{t-1} are small K-major B tiles in shared
// memory.  Precision as in the convolutions: every operand is a bf16 pair v = hi + lo; the B tile
// stacks hi and lo along N (columns 0..31 clips' hi, 32..63 clips' lo) and the weights' hi and lo
// parts are two A operands, so one accumulator holds all four partial products and the epilogue
// adds column n and 32 + n.
//
// Gate rows are permuted so that M-tile m, lane 32*g + u is gate type g (i,f,g,o) of unit 32m + u:
// TMEM lane quadrant q == gate type q.  8 epilogue warps: warp (q, m) reads its quadrant of tile m,
// applies e = exp(-v) (exp(-2v) for the cell candidate) and parks e in shared memory; after a
// named barrier, thread (chunk c = warp, clip = lane) owns 8 units of one clip, keeps their cell
// state in registers and evaluates, with ONE reciprocal each,
//     c' = (c (1+e_i)(1+e_g) + (1-e_g)(1+e_f)) / ((1+e_f)(1+e_i)(1+e_g))     [= s(f) c + s(i) tanh(g)]
//     h  = (1-e_c) / ((1+e_o)(1+e_c)),  e_c = exp(-2c')                       [= s(o) tanh(c')]
// then writes h (bf16 hi/lo) into the h tile for the next step's MMA and into the planar output.
// x_t arrives through 4 loader warps (cp.async, transposing planar [clip][plane][t] into per-step
// [plane][clip] tiles, 8 steps per stage, double buffered).
#include <cuda_bf16.h>

#include "wm_common.h"
#include "wm_tc.cuh"

namespace wm {

using namespace tc;

namespace {

constexpr int PAD = WM_PLANAR_PAD;
constexpr int NCL = 32;                       // clips per CTA
constexpr int TC_STEPS = 8;                   // steps per x stage
constexpr int XTILE = 16 * NCL * 16;          // bytes of one step's x tile: [chunk 8][hi/lo 2][clip 32][16 B]
constexpr int XSTEP = XTILE + 16;             // padded pitch between steps (bank spread for the transposing stores)
constexpr int XSTAGE = TC_STEPS * XSTEP;
constexpr int HTILE = XTILE;                  // h tile, same layout
constexpr int ELD = NCL + 1;                  // padded row of the exchange buffer
constexpr int EBYTES = 4 * 64 * ELD * 4;      // e[gate][unit][clip]
constexpr int OFF_X = 0;
constexpr int OFF_H = OFF_X + 2 * XSTAGE;
constexpr int OFF_E = OFF_H + HTILE;
constexpr int OFF_BAR = OFF_E + EBYTES;
constexpr int LSTM_SMEM = OFF_BAR + 128;
constexpr int N_EPI = 256, N_LOAD = 128;
constexpr int THREADS = N_EPI + N_LOAD + 32;
constexpr uint32_t kIdescL = make_idesc(128, 64);
// TMEM columns: weights [mat 4][tile 2] x 32 columns, then accumulators [buf 2][tile 2] x 64 columns
constexpr uint32_t TM_W = 0, TM_ACC = 256;
constexpr float kLog2e = 1.4426950408889634f;

}  // namespace

// wpk: bf16 [mat: Whh_hi, Whh_lo, Wih_hi, Wih_lo][tile 2][lane 128][64]; bias_p: fp32 [tile 2][lane 128]
__global__ void __launch_bounds__(THREADS, 1)
    lstm_tc_kernel(const uint4 *__restrict__ x, const uint4 *__restrict__ wpk, const float *__restrict__ bias_p,
                   const float *__restrict__ chan_add, uint4 *__restrict__ y, int B, int T,
                   long long *__restrict__ prof) {
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t s_base = smem_u32(smem);
  const uint32_t bars = s_base + OFF_BAR;
  const uint32_t h_ready = bars, acc_full0 = bars + 8, acc_empty0 = bars + 24, x_full0 = bars + 40,
                 x_empty0 = bars + 56;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + OFF_BAR + 96);
  float *E = reinterpret_cast<float *>(smem + OFF_E);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b0 = blockIdx.x * NCL;
  const int nb = min(NCL, B - b0);
  const size_t RP = (size_t)T + 2 * PAD;
  const int nchunk = (T + TC_STEPS - 1) / TC_STEPS;

  if (tid == 0) {
    mbar_init(h_ready, N_EPI / 32);
    for (int i = 0; i < 2; ++i) {
      mbar_init(acc_full0 + 8 * i, 1);
      mbar_init(acc_empty0 + 8 * i, N_EPI / 32);
      mbar_init(x_full0 + 8 * i, N_LOAD / 32);
      mbar_init(x_empty0 + 8 * i, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // h_0 = 0 and clean x stages (columns of absent clips must at least be finite)
  for (int i = tid; i < (OFF_E) / 16; i += THREADS) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (warp == 12) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp < 4) {
    // weights -> TMEM: lane = gate row, 32 columns = 64 bf16 (two per 32-bit column, K ascending)
    const int q = warp;
#pragma unroll 1
    for (int mt = 0; mt < 8; ++mt) {
      const uint4 *src = wpk + ((size_t)mt * 128 + q * 32 + lane) * 8;
      uint32_t r[32];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        uint4 v = __ldg(src + i);
        r[4 * i] = v.x; r[4 * i + 1] = v.y; r[4 * i + 2] = v.z; r[4 * i + 3] = v.w;
      }
      tmem_st32(tmem + TM_W + mt * 32 + ((uint32_t)(q * 32) << 16), r);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp < 8) {
    // ===================== epilogue / cell update =====================
    const int q = warp & 3, m = warp >> 2;                 // phase 1: gate type q, M-tile m
    const float bias = bias_p[m * 128 + q * 32 + lane];
    const float escale = (q == 2) ? -2.0f * kLog2e : -kLog2e;
    const float vmax = (q == 2) ? 10.0f : 20.0f;
    float *e_row = E + (q * 64 + m * 32 + lane) * ELD;      // phase 1 writes e[q][unit][0..31]
    const int c8 = warp, n = lane;                          // phase 2: units 8*c8.., clip n
    const bool live = n < nb;
    float cst[8], emb[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      cst[i] = 0.0f;
      emb[i] = (chan_add != nullptr && live) ? chan_add[(size_t)(b0 + n) * 64 + c8 * 8 + i] : 0.0f;
    }
    uint4 *y_hi = y + ((size_t)(b0 + n) * 16 + c8) * RP + PAD;
    uint4 *y_lo = y + ((size_t)(b0 + n) * 16 + 8 + c8) * RP + PAD;
    if (live) {  // the planes' zero padding rows
      for (int r = 0; r < PAD; ++r) {
        y_hi[-1 - r] = make_uint4(0, 0, 0, 0); y_lo[-1 - r] = make_uint4(0, 0, 0, 0);
        y_hi[T + r] = make_uint4(0, 0, 0, 0); y_lo[T + r] = make_uint4(0, 0, 0, 0);
      }
    }
    uint4 *h_hi = reinterpret_cast<uint4 *>(smem + OFF_H + c8 * (2 * NCL * 16) + n * 16);
    uint4 *h_lo = h_hi + NCL;
    const bool pf = prof != nullptr && blockIdx.x == 0 && tid == 0;
    long long pa[7] = {0, 0, 0, 0, 0, 0, 0};
#pragma unroll 1
    for (int t = 0; t < T; ++t) {
      const int buf = t & 1;
      long long c0 = pf ? clock64() : 0;
      mbar_wait_warp(acc_full0 + 8 * buf, (t >> 1) & 1);
      tc_fence_after();
      long long c1 = pf ? clock64() : 0;
      uint32_t r[64];
      const uint32_t ta = tmem + TM_ACC + buf * 128 + m * 64 + ((uint32_t)(q * 32) << 16);
      tmem_ld32(ta, r);
      tmem_ld32(ta + 32, r + 32);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive_warp(acc_empty0 + 8 * buf);
      long long c2 = pf ? clock64() : 0;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float v = __uint_as_float(r[j]) + __uint_as_float(r[32 + j]) + bias;
        v = fminf(fmaxf(v, -vmax), vmax);
        e_row[j] = ex2_approx(v * escale);
      }
      long long c3 = pf ? clock64() : 0;
      named_bar_sync(1, N_EPI);
      long long c4 = pf ? clock64() : 0;
      float hv[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int u = c8 * 8 + i;
        const float ei = E[(0 * 64 + u) * ELD + n], ef = E[(1 * 64 + u) * ELD + n];
        const float eg = E[(2 * 64 + u) * ELD + n], eo = E[(3 * 64 + u) * ELD + n];
        const float ag = (1.0f + ei) * (1.0f + eg), bf = 1.0f + ef;
        const float cn = (cst[i] * ag + (1.0f - eg) * bf) * rcp_approx(bf * ag);
        cst[i] = cn;
        const float ec = ex2_approx(fminf(fmaxf(cn, -10.0f), 10.0f) * (-2.0f * kLog2e));
        hv[i] = (1.0f - ec) * rcp_approx((1.0f + eo) * (1.0f + ec));
      }
      uint4 hi, lo;
      split8(hv, hi, lo);
      *h_hi = hi;
      *h_lo = lo;
      long long c5 = pf ? clock64() : 0;
      fence_async_smem();
      mbar_arrive_warp(h_ready);
      long long c6 = pf ? clock64() : 0;
      if (live) {
        if (chan_add != nullptr) {
#pragma unroll
          for (int i = 0; i < 8; ++i) hv[i] += emb[i];
          split8(hv, hi, lo);
        }
        y_hi[t] = hi;
        y_lo[t] = lo;
      }
      if (pf) {
        long long c7 = clock64();
        pa[0] += c1 - c0; pa[1] += c2 - c1; pa[2] += c3 - c2; pa[3] += c4 - c3; pa[4] += c5 - c4; pa[5] += c6 - c5;
        pa[6] += c7 - c6;
      }
    }
    if (pf) {
      for (int i = 0; i < 7; ++i) prof[i] = pa[i];
    }
  } else if (warp < 12) {
    // ===================== x loader =====================
    const int lt = tid - N_EPI;                 // 0..127: plane = lt / 8, step-in-stage = lt % 8
    const int pl = lt >> 3, tt = lt & 7;
    const uint32_t dst0 = s_base + OFF_X + tt * XSTEP + (pl & 7) * (2 * NCL * 16) + (pl >> 3) * (NCL * 16);
#pragma unroll 1
    for (int ch = 0; ch < nchunk; ++ch) {
      const int st = ch & 1;
      if (ch >= 2) mbar_wait_warp(x_empty0 + 8 * st, ((ch >> 1) - 1) & 1);
      const int t = ch * TC_STEPS + tt;
      if (t < T) {
        const uint4 *src = x + ((size_t)b0 * 16 + pl) * RP + PAD + t;
        const uint32_t dst = dst0 + st * XSTAGE;
        for (int nn = 0; nn < nb; ++nn) cp_async16(dst + nn * 16, src + (size_t)nn * 16 * RP);
      }
      cp_async_commit();
      if (ch >= 1) {  // the previous stage has landed: publish it
        cp_async_wait<1>();
        fence_async_smem();
        mbar_arrive_warp(x_full0 + 8 * ((ch - 1) & 1));
      }
    }
    cp_async_wait<0>();
    fence_async_smem();
    mbar_arrive_warp(x_full0 + 8 * ((nchunk - 1) & 1));
  } else {
    // ===================== MMA issuer (whole warp walks the pipeline; one elected lane issues) =====
    const bool issuer = elect_one();
    const uint64_t h_desc = smem_desc(s_base + OFF_H, 2 * NCL * 16, 128);
    // 2 M-tiles x 4 K steps x {hi, lo} weights; B tile: 64 rows (32 hi + 32 lo clips), chunk pitch 1 KB
    auto issue = [&](uint64_t b_desc, uint32_t w_first, int buf, uint32_t accum) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
        for (int part = 0; part < 2; ++part) {
#pragma unroll
          for (int m = 0; m < 2; ++m) {
            mma_bf16_ts(tmem + TM_ACC + buf * 128 + m * 64, tmem + TM_W + (w_first + part) * 64 + m * 32 + kk * 8,
                        b_desc + (uint64_t)(((2 * kk) * (2 * NCL * 16)) >> 4), kIdescL, (kk | part) != 0 ? 1u : accum);
          }
        }
      }
    };
    const bool pf = prof != nullptr && blockIdx.x == 0 && issuer;
    long long pm[3] = {0, 0, 0};
    // x part of step 0
    mbar_wait_warp(x_full0, 0);
    tc_fence_after();
    if (issuer) issue(smem_desc(s_base + OFF_X, 2 * NCL * 16, 128), 2, 0, 0);
    __syncwarp();
#pragma unroll 1
    for (int t = 0; t < T; ++t) {
      const int buf = t & 1;
      long long m0 = pf ? clock64() : 0;
      if (t > 0) {
        mbar_wait_warp(h_ready, (t - 1) & 1);
        tc_fence_after();
      }
      long long m1 = pf ? clock64() : 0;
      if (issuer) {
        issue(h_desc, 0, buf, 1);                 // + W_hh . h_{t-1}
        tc_commit(acc_full0 + 8 * buf);
      }
      __syncwarp();
      long long m2 = pf ? clock64() : 0;
      const int t1 = t + 1;
      if (t1 < T) {
        const int ch = t1 / TC_STEPS, tt = t1 % TC_STEPS, st = ch & 1;
        if (tt == 0) mbar_wait_warp(x_full0 + 8 * st, (ch >> 1) & 1);
        if (t1 >= 2) mbar_wait_warp(acc_empty0 + 8 * (buf ^ 1), ((t1 >> 1) - 1) & 1);
        tc_fence_after();
        if (issuer) {
          issue(smem_desc(s_base + OFF_X + st * XSTAGE + tt * XSTEP, 2 * NCL * 16, 128), 2, buf ^ 1, 0);  // W_ih . x_{t+1}
          if (tt == TC_STEPS - 1) tc_commit(x_empty0 + 8 * st);
        }
        __syncwarp();
      }
      if (pf) { long long m3 = clock64(); pm[0] += m1 - m0; pm[1] += m2 - m1; pm[2] += m3 - m2; }
    }
    if (pf) {
      for (int i = 0; i < 3; ++i) prof[8 + i] = pm[i];
    }
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 12) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
  }
}

// optional device buffer of 32 int64 receiving per-phase cycle sums of block 0 (tools/*_profile.py)
static long long *g_lstm_prof = nullptr;
void set_lstm_profile_buffer(long long *p) { g_lstm_prof = p; }
long long *get_profile_buffer() { return g_lstm_prof; }

int launch_lstm_tc(const void *x, const void *wpk, const float *bias_p, const float *chan_add, void *y, int B, int T,
                   cudaStream_t st) {
  if (B == 0 || T == 0) return 0;
  static bool attr_set = false;
  if (!attr_set) {
    WM_CHECK_CUDA(cudaFuncSetAttribute(lstm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LSTM_SMEM));
    attr_set = true;
  }
  lstm_tc_kernel<<<(B + NCL - 1) / NCL, THREADS, LSTM_SMEM, st>>>(
      reinterpret_cast<const uint4 *>(x), reinterpret_cast<const uint4 *>(wpk), bias_p, chan_add,
      reinterpret_cast<uint4 *>(y), B, T, g_lstm_prof);
  WM_CHECK_LAUNCH("lstm_tc");
  return 0;
}

// fp32 W_ih, W_hh [256][64] (rows i,f,g,o) and bias[256] -> the kernel's packed operands:
//   wpk  bf16 [mat: Whh_hi, Whh_lo, Wih_hi, Wih_lo][tile 2][lane 128][64],  row (tile m, lane l) = gate l/32 of unit 32m + l%32
//   bias_p fp32 [tile 2][lane 128]
__global__ void pack_lstm_tc_kernel(const float *__restrict__ w_ih, const float *__restrict__ w_hh,
                                    const float *__restrict__ bias, __nv_bfloat16 *__restrict__ wpk,
                                    float *__restrict__ bias_p) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < 256) {
    int m = e >> 7, l = e & 127;
    bias_p[e] = bias[(l >> 5) * 64 + m * 32 + (l & 31)];
  }
  if (e >= 4 * 256 * 64) return;
  int k = e & 63, l = (e >> 6) & 127, m = (e >> 13) & 1, mat = e >> 14;
  int row = (l >> 5) * 64 + m * 32 + (l & 31);
  float v = (mat < 2 ? w_hh : w_ih)[row * 64 + k];
  __nv_bfloat16 hi = __float2bfloat16_rn(v);
  wpk[e] = (mat & 1) ? __float2bfloat16_rn(v - __bfloat162float(hi)) : hi;
}

int launch_pack_lstm_tc(const float *w_ih, const float *w_hh, const float *bias, void *wpk, float *bias_p,
                        cudaStream_t st) {
  pack_lstm_tc_kernel<<<(4 * 256 * 64 + 255) / 256, 256, 0, st>>>(w_ih, w_hh, bias,
                                                                 reinterpret_cast<__nv_bfloat16 *>(wpk), bias_p);
  WM_CHECK_LAUNCH("pack_lstm_tc");
  return 0;
}

}  // namespace wm
