// nn.LSTM(64,64,batch_first=True), zero initial state (py/main16.py:138,153) as a persistent
// tensor-core kernel (sm_100a).  One CTA owns 32 clips for all T steps, as TWO independent groups of 16.
//
// Per step the gate pre-activations are computed TRANSPOSED,
//     G^T[256 gate rows, clips] = W_ih . x_t^T + W_hh . h_{t-1}^T ,
// so the (constant) weights are the A operand and stay RESIDENT IN TENSOR MEMORY for the whole
// kernel (tcgen05.mma with A in TMEM), while x_t / h_{t-1} are small K-major B tiles in shared
// memory.  Precision as in the convolutions: every operand is a bf16 pair v = hi + lo; the B tile
// stacks hi and lo along N (rows 0..15 the clips' hi, 16..31 their lo); W_hi multiplies all 32 rows,
// W_lo only the hi rows (the lo x lo product is below the parity budget), and the epilogue adds
// accumulator column n and 16 + n.
//
// The step is a dependent chain (MMA -> TMEM load -> exp -> cell update -> h tile -> proxy fence -> MMA)
// that leaves the tensor pipe and the ALUs idle most of the time, so the CTA runs two such chains,
// half a step apart: each group has its own 8 epilogue warps, 2 loader warps, MMA-issuing warp,
// barriers, accumulators and tiles; only the weights in TMEM are shared.
//
// Gate rows are permuted so that M-tile m, lane 32*g + u is gate type g (i,f,g,o) of unit 32m + u:
// TMEM lane quadrant q == gate type q.  The packed weights and biases are pre-multiplied by -log2(e)
// (-2 log2(e) for the cell candidate), so the accumulator IS the ex2 argument.  Epilogue warp (q, m)
// reads its quadrant of tile m, e = ex2(min(v, 28.85)), and parks e in shared memory; after a named
// barrier, thread (8-unit chunk = warp, half = lane / 16, clip = lane % 16) owns 4 units of one clip,
// keeps their cell state in registers and evaluates, with ONE reciprocal each,
//     c' = (c (1+e_i)(1+e_g) + (1-e_g)(1+e_f)) / ((1+e_f)(1+e_i)(1+e_g))     [= s(f) c + s(i) tanh(g)]
//     h  = (1-e_c) / ((1+e_o)(1+e_c)),  e_c = exp(-2c')                       [= s(o) tanh(c')]
// The two half-lanes of a clip swap their bf16 hi / lo halves so that each writes one 16-byte row of
// the h tile (the next step's B operand) and of the planar output.
// x_t arrives through the loader warps (cp.async, transposing planar [clip][plane][t] into per-step
// [plane][clip] tiles, 8 steps per stage, double buffered).
#include <cuda_bf16.h>

#include "wm_common.h"
#include "wm_tc.cuh"

namespace wm {

using namespace tc;

namespace {

constexpr int PAD = WM_PLANAR_PAD;
constexpr int NCL = 16;                       // clips per group
constexpr int NGRP = 2;                       // groups per CTA
constexpr int TC_STEPS = 8;                   // steps per x stage
constexpr int XTILE = 16 * NCL * 16;          // bytes of one step's x tile: [chunk 8][hi/lo 2][clip 16][16 B]
constexpr int XSTEP = XTILE + 16;             // padded pitch between steps (bank spread for the transposing stores)
constexpr int XSTAGE = TC_STEPS * XSTEP;
constexpr int HTILE = XTILE;                  // h tile, same layout
constexpr int ELD = NCL + 4;                  // row pitch of the exchange buffer: 16-byte rows; 20 words spreads both the
                                              // phase-1 STS.128 (rows = lanes) and the phase-2 reads (half-lanes 4 units apart) over all banks
constexpr int EBYTES = 4 * 64 * ELD * 4;      // e[gate][unit][clip]
constexpr int G_X = 0;                        // offsets inside one group's shared memory
constexpr int G_H = G_X + 2 * XSTAGE;
constexpr int G_E = G_H + HTILE;
constexpr int G_BYTES = ((G_E + EBYTES + 127) / 128) * 128;
constexpr int OFF_BAR = NGRP * G_BYTES;
constexpr int LSTM_SMEM = OFF_BAR + 256;
constexpr int N_EPI = 256, N_LOAD = 64;       // per group
constexpr int W_LOAD0 = NGRP * N_EPI / 32, W_MMA0 = W_LOAD0 + NGRP * N_LOAD / 32;
constexpr int THREADS = NGRP * (N_EPI + N_LOAD + 32);
constexpr uint32_t kIdescHi = make_idesc(128, 2 * NCL);   // W_hi x [clips hi | clips lo]
constexpr uint32_t kIdescLo = make_idesc(128, NCL);       // W_lo x  clips hi
// TMEM columns: weights [mat 4][tile 2] x 32 columns, then accumulators [group 2][buf 2][tile 2] x 32 columns
constexpr uint32_t TM_W = 0, TM_ACC = 256;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kArgMax = 28.85f;             // ex2 argument cap: e <= 4.8e8, (1+e)^3 stays finite

}  // namespace

// wpk: bf16 [mat: Whh_hi, Whh_lo, Wih_hi, Wih_lo][tile 2][lane 128][64]; bias_p: fp32 [tile 2][lane 128]
// (both pre-scaled by -log2 e / -2 log2 e, see launch_pack_lstm_tc)
template <bool PROF>
__global__ void __launch_bounds__(THREADS, 1)
    lstm_tc_kernel(const uint4 *__restrict__ x, const uint4 *__restrict__ wpk, const float *__restrict__ bias_p,
                   const float *__restrict__ chan_add, uint4 *__restrict__ y, int B, int T,
                   long long *__restrict__ prof) {
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t s_base = smem_u32(smem);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + OFF_BAR + 2 * 96);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const size_t RP = (size_t)T + 2 * PAD;
  const int nchunk = (T + TC_STEPS - 1) / TC_STEPS;

  if (tid == 0) {
    for (int g = 0; g < NGRP; ++g) {
      const uint32_t bars = s_base + OFF_BAR + 96 * g;
      mbar_init(bars, N_EPI / 32);                       // h_ready
      for (int i = 0; i < 2; ++i) {
        mbar_init(bars + 8 + 8 * i, 1);                  // acc_full
        mbar_init(bars + 24 + 8 * i, N_EPI / 32);        // acc_empty
        mbar_init(bars + 40 + 8 * i, N_LOAD / 32);       // x_full
        mbar_init(bars + 56 + 8 * i, 1);                 // x_empty
      }
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // h_0 = 0 and clean x stages (columns of absent clips must at least be finite)
  for (int i = tid; i < OFF_BAR / 16; i += THREADS) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (warp == W_MMA0) {
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

  // ---- role and group of this warp ----
  int role, g;                                  // role 0 epilogue, 1 loader, 2 MMA issuer
  if (warp < W_LOAD0) { role = 0; g = warp / (N_EPI / 32); }
  else if (warp < W_MMA0) { role = 1; g = (warp - W_LOAD0) / (N_LOAD / 32); }
  else { role = 2; g = warp - W_MMA0; }
  const int b0 = blockIdx.x * (NGRP * NCL) + g * NCL;      // first clip of the group
  const int nb = max(0, min(NCL, B - b0));
  const uint32_t gs = s_base + g * G_BYTES;
  uint8_t *gsm = smem + g * G_BYTES;
  const uint32_t bars = s_base + OFF_BAR + 96 * g;
  const uint32_t h_ready = bars, acc_full0 = bars + 8, acc_empty0 = bars + 24, x_full0 = bars + 40,
                 x_empty0 = bars + 56;
  const uint32_t tm_acc = tmem + TM_ACC + g * 128;

  if (nb > 0 && role == 0) {
    // ===================== epilogue / cell update =====================
    const int w = warp & 7;
    const int q = w & 3, m = w >> 2;                        // phase 1: gate type q, M-tile m
    const float bias = bias_p[m * 128 + q * 32 + lane];
    float *E = reinterpret_cast<float *>(gsm + G_E);
    float *e_row = E + (q * 64 + m * 32 + lane) * ELD;      // phase 1 writes e[q][unit][0..15]
    const int c8 = w, hh = lane >> 4, n = lane & 15;        // phase 2: units 8*c8 + 4*hh .., clip n
    const int u0 = c8 * 8 + hh * 4;
    const bool live = n < nb;
    f32x2 cst[2];                                           // cell state, scaled by -2 log2(e), units (u0,u0+1), (u0+2,u0+3)
    float emb[4];
    cst[0] = cst[1] = pk2(0.0f, 0.0f);
#pragma unroll
    for (int i = 0; i < 4; ++i)
      emb[i] = (chan_add != nullptr && live) ? chan_add[(size_t)(b0 + n) * 64 + u0 + i] : 0.0f;
    // this lane stores one 16-byte row per step: the hi plane (hh = 0) or the lo plane (hh = 1) of chunk c8
    uint4 *y_row = y + ((size_t)(b0 + n) * 16 + hh * 8 + c8) * RP + PAD;
    if (live) {  // the planes' zero padding rows
      for (int r = 0; r < PAD; ++r) { y_row[-1 - r] = make_uint4(0, 0, 0, 0); y_row[T + r] = make_uint4(0, 0, 0, 0); }
    }
    uint4 *h_row = reinterpret_cast<uint4 *>(gsm + G_H + c8 * (2 * NCL * 16) + hh * (NCL * 16) + n * 16);
    const bool pf = PROF && prof != nullptr && blockIdx.x == 0 && tid == 0;
    long long pa[7] = {0, 0, 0, 0, 0, 0, 0};
    const f32x2 bias2 = pk2(bias, bias), one2 = pk2(1.0f, 1.0f), mone2 = pk2(-1.0f, -1.0f);
    const f32x2 k2 = pk2(-2.0f * kLog2e, -2.0f * kLog2e), mk2 = pk2(2.0f * kLog2e, 2.0f * kLog2e);
    const float *Eu = E + u0 * ELD + n;
#pragma unroll 1
    for (int t = 0; t < T; ++t) {
      const int buf = t & 1;
      long long c0 = pf ? clock64() : 0;
      mbar_wait_warp(acc_full0 + 8 * buf, (t >> 1) & 1);
      tc_fence_after();
      long long c1 = pf ? clock64() : 0;
      uint32_t r[32];
      tmem_ld32(tm_acc + buf * 64 + m * 32 + ((uint32_t)(q * 32) << 16), r);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive_warp(acc_empty0 + 8 * buf);
      long long c2 = pf ? clock64() : 0;
      // phase 1: e = 2^min(acc_hi + acc_lo + bias, cap) for this gate row and the 16 clips
#pragma unroll
      for (int j4 = 0; j4 < NCL / 4; ++j4) {
        float e[4];
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
          const int j = j4 * 4 + h2 * 2;
          const f32x2 v = add2(add2(pk2(__uint_as_float(r[j]), __uint_as_float(r[j + 1])),
                                    pk2(__uint_as_float(r[NCL + j]), __uint_as_float(r[NCL + j + 1]))), bias2);
          float va, vb;
          upk2(v, va, vb);
          e[h2 * 2] = ex2_approx(fminf(va, kArgMax));
          e[h2 * 2 + 1] = ex2_approx(fminf(vb, kArgMax));
        }
        *reinterpret_cast<float4 *>(e_row + j4 * 4) = make_float4(e[0], e[1], e[2], e[3]);
      }
      long long c3 = pf ? clock64() : 0;
      named_bar_sync(1 + g, N_EPI);
      long long c4 = pf ? clock64() : 0;
      // phase 2: cell update of units u0..u0+3 of clip n, two units per packed instruction
      float hv[4];
#pragma unroll
      for (int p2 = 0; p2 < 2; ++p2) {
        const float *Ep = Eu + 2 * p2 * ELD;
        const f32x2 ei = pk2(Ep[0], Ep[ELD]), ef = pk2(Ep[64 * ELD], Ep[65 * ELD]);
        const f32x2 eg = pk2(Ep[128 * ELD], Ep[129 * ELD]), eo = pk2(Ep[192 * ELD], Ep[193 * ELD]);
        const f32x2 ag = mul2(add2(ei, one2), add2(eg, one2)), bf = add2(ef, one2);
        const f32x2 kg = fma2(eg, mk2, k2);                      // -2 log2(e) (1 - e_g)
        const f32x2 num = fma2(kg, bf, mul2(cst[p2], ag));
        float da, db;
        upk2(mul2(bf, ag), da, db);
        const f32x2 cn = mul2(num, pk2(rcp_approx(da), rcp_approx(db)));
        cst[p2] = cn;
        float ca, cb;
        upk2(cn, ca, cb);
        const f32x2 ec = pk2(ex2_approx(fminf(ca, kArgMax)), ex2_approx(fminf(cb, kArgMax)));   // exp(-2 c')
        float ha, hb;
        upk2(mul2(add2(eo, one2), add2(ec, one2)), ha, hb);
        upk2(mul2(fma2(ec, mone2, one2), pk2(rcp_approx(ha), rcp_approx(hb))), hv[2 * p2], hv[2 * p2 + 1]);
      }
      uint2 hi, lo;
      split4(hv, hi, lo);
      {  // half-lane 0 collects the 8 hi values of the chunk, half-lane 1 the 8 lo values
        const uint2 snd = hh ? hi : lo;
        const uint32_t rx = __shfl_xor_sync(0xffffffffu, snd.x, 16), ry = __shfl_xor_sync(0xffffffffu, snd.y, 16);
        *h_row = hh ? make_uint4(rx, ry, lo.x, lo.y) : make_uint4(hi.x, hi.y, rx, ry);
      }
      long long c5 = pf ? clock64() : 0;
      fence_async_smem();
      // hand h_t to the MMA warp through a hardware named barrier (arrive here, sync there): cheaper than an
      // mbarrier round trip on the step's critical path
      asm volatile("bar.arrive %0, %1;" ::"r"(3 + g), "r"(N_EPI + 32) : "memory");
      long long c6 = pf ? clock64() : 0;
      uint4 out;
      if (chan_add != nullptr) {   // warp-uniform
#pragma unroll
        for (int i = 0; i < 4; ++i) hv[i] += emb[i];
        split4(hv, hi, lo);
      }
      {
        const uint2 snd = hh ? hi : lo;
        const uint32_t rx = __shfl_xor_sync(0xffffffffu, snd.x, 16), ry = __shfl_xor_sync(0xffffffffu, snd.y, 16);
        out = hh ? make_uint4(rx, ry, lo.x, lo.y) : make_uint4(hi.x, hi.y, rx, ry);
      }
      if (live) y_row[t] = out;
      if (pf) {
        long long c7 = clock64();
        pa[0] += c1 - c0; pa[1] += c2 - c1; pa[2] += c3 - c2; pa[3] += c4 - c3; pa[4] += c5 - c4; pa[5] += c6 - c5;
        pa[6] += c7 - c6;
      }
    }
    if (pf) {
      for (int i = 0; i < 7; ++i) prof[i] = pa[i];
    }
  } else if (nb > 0 && role == 1) {
    // ===================== x loader =====================
    const int lt = tid - (W_LOAD0 * 32 + g * N_LOAD);   // 0..63: two (plane, step-in-stage) pairs each
#pragma unroll 1
    for (int ch = 0; ch < nchunk; ++ch) {
      const int st = ch & 1;
      if (ch >= 2) mbar_wait_warp(x_empty0 + 8 * st, ((ch >> 1) - 1) & 1);
#pragma unroll
      for (int rep = 0; rep < 2; ++rep) {
        const int pr = lt + rep * N_LOAD;
        const int pl = pr >> 3, tt = pr & 7;
        const int t = ch * TC_STEPS + tt;
        if (t < T) {
          const uint4 *src = x + ((size_t)b0 * 16 + pl) * RP + PAD + t;
          const uint32_t dst = gs + G_X + st * XSTAGE + tt * XSTEP + (pl & 7) * (2 * NCL * 16) + (pl >> 3) * (NCL * 16);
          for (int nn = 0; nn < nb; ++nn) cp_async16(dst + nn * 16, src + (size_t)nn * 16 * RP);
        }
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
  } else if (nb > 0 && role == 2) {
    // ===================== MMA issuer (whole warp walks the pipeline; one elected lane issues) =====
    const bool issuer = elect_one();
    const uint64_t h_desc = smem_desc(gs + G_H, 2 * NCL * 16, 128);
    // 2 M-tiles x 4 K steps x {W_hi (N = 32), W_lo (N = 16)}; B tile: 32 rows (16 hi + 16 lo clips), chunk pitch 512 B
    auto issue = [&](uint64_t b_desc, uint32_t w_first, int buf, uint32_t accum) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
        for (int part = 0; part < 2; ++part) {
#pragma unroll
          for (int m = 0; m < 2; ++m) {
            mma_bf16_ts(tm_acc + buf * 64 + m * 32, tmem + TM_W + (w_first + part) * 64 + m * 32 + kk * 8,
                        b_desc + (uint64_t)(((2 * kk) * (2 * NCL * 16)) >> 4), part ? kIdescLo : kIdescHi,
                        (kk | part) != 0 ? 1u : accum);
          }
        }
      }
    };
    const bool pf = PROF && prof != nullptr && blockIdx.x == 0 && g == 0 && issuer;
    long long pm[3] = {0, 0, 0};
    // x part of step 0
    mbar_wait_warp(x_full0, 0);
    tc_fence_after();
    if (issuer) issue(smem_desc(gs + G_X, 2 * NCL * 16, 128), 2, 0, 0);
    __syncwarp();
#pragma unroll 1
    for (int t = 0; t < T; ++t) {
      const int buf = t & 1;
      long long m0 = pf ? clock64() : 0;
      if (t > 0) {
        asm volatile("bar.sync %0, %1;" ::"r"(3 + g), "r"(N_EPI + 32) : "memory");
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
          issue(smem_desc(gs + G_X + st * XSTAGE + tt * XSTEP, 2 * NCL * 16, 128), 2, buf ^ 1, 0);  // W_ih . x_{t+1}
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
  if (warp == W_MMA0) {
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
    WM_CHECK_CUDA(cudaFuncSetAttribute(lstm_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, LSTM_SMEM));
    WM_CHECK_CUDA(cudaFuncSetAttribute(lstm_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, LSTM_SMEM));
    attr_set = true;
  }
  const int grid = (B + NGRP * NCL - 1) / (NGRP * NCL);
  if (g_lstm_prof != nullptr)   // developer build of the same kernel with per-phase cycle counters (tools/lstm_profile.py)
    lstm_tc_kernel<true><<<grid, THREADS, LSTM_SMEM, st>>>(
        reinterpret_cast<const uint4 *>(x), reinterpret_cast<const uint4 *>(wpk), bias_p, chan_add,
        reinterpret_cast<uint4 *>(y), B, T, g_lstm_prof);
  else
    lstm_tc_kernel<false><<<grid, THREADS, LSTM_SMEM, st>>>(
        reinterpret_cast<const uint4 *>(x), reinterpret_cast<const uint4 *>(wpk), bias_p, chan_add,
        reinterpret_cast<uint4 *>(y), B, T, nullptr);
  WM_CHECK_LAUNCH("lstm_tc");
  return 0;
}

// fp32 W_ih, W_hh [256][64] (rows i,f,g,o) and bias[256] -> the kernel's packed operands:
//   wpk  bf16 [mat: Whh_hi, Whh_lo, Wih_hi, Wih_lo][tile 2][lane 128][64],  row (tile m, lane l) = gate l/32 of unit 32m + l%32
//   bias_p fp32 [tile 2][lane 128]
// Rows and biases are multiplied by -log2(e) (gates i, f, o) / -2 log2(e) (cell candidate g) before the
// hi/lo split, so the kernel's accumulator is directly the argument of ex2 in sigmoid / tanh.
__global__ void pack_lstm_tc_kernel(const float *__restrict__ w_ih, const float *__restrict__ w_hh,
                                    const float *__restrict__ bias, __nv_bfloat16 *__restrict__ wpk,
                                    float *__restrict__ bias_p) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < 256) {
    int m = e >> 7, l = e & 127;
    bias_p[e] = bias[(l >> 5) * 64 + m * 32 + (l & 31)] * ((l >> 5) == 2 ? -2.0f * kLog2e : -kLog2e);
  }
  if (e >= 4 * 256 * 64) return;
  int k = e & 63, l = (e >> 6) & 127, m = (e >> 13) & 1, mat = e >> 14;
  int row = (l >> 5) * 64 + m * 32 + (l & 31);
  float v = (mat < 2 ? w_hh : w_ih)[row * 64 + k] * ((l >> 5) == 2 ? -2.0f * kLog2e : -kLog2e);
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
