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
// The step is a dependent chain (MMA -> TMEM load -> exp -> cell update -> h tile -> MMA) that leaves the
// tensor pipe and the ALUs idle most of the time, so the CTA runs two such chains: each group has its
// own 8 epilogue warps, MMA-issuing warp, barriers, accumulators and tiles; only the weights in TMEM are shared.
//
// Waiting is done on HARDWARE named barriers wherever a whole warp waits: a warp spinning on an mbarrier
// keeps taking issue slots from the other group's warps on its scheduler (measured: the exponentials of one
// group ran at half speed while the other group's 8 warps polled).  Only the elected MMA thread polls
// mbarriers (accumulator complete, x stage landed / consumed) and relays "accumulator ready" to its epilogue
// warps with bar.arrive.  The generic->async proxy fence for the h tile is issued by that thread too (after
// the barrier that orders the epilogue warps' stores), not by the 256 epilogue threads on the critical path.
//
// x_t arrives by TMA: one 5-d tensor map over the planar input (16 B row, clip, hi/lo, 8-channel chunk, time)
// whose box {16 B, 16 clips, 2, 8, 8 steps} lands in shared memory already transposed into 8 per-step
// [chunk][hi/lo][clip] B tiles (UTMALDG); clips beyond B are zero-filled by the out-of-bounds rule.
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
#include <cuda.h>
#include <cuda_bf16.h>
#include <cudaTypedefs.h>

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
constexpr int XSTAGE = TC_STEPS * XTILE;      // one TMA box
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
constexpr int N_EPI = 256;                    // epilogue threads per group
constexpr int W_MMA0 = NGRP * N_EPI / 32;
constexpr int THREADS = NGRP * (N_EPI + 32);
constexpr uint32_t kIdescHi = make_idesc(128, 2 * NCL);   // W_hi x [clips hi | clips lo]
constexpr uint32_t kIdescLo = make_idesc(128, NCL);       // W_lo x  clips hi
// TMEM columns: weights [mat 4][tile 2] x 32 columns, then accumulators [group 2][buf 2][tile 2] x 32 columns
constexpr uint32_t TM_W = 0, TM_ACC = 256;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kArgMax = 28.85f;             // ex2 argument cap: e <= 4.8e8, (1+e)^3 stays finite
// hardware named barriers (0 = __syncthreads): per group g
constexpr int BAR_PHASE = 1, BAR_HREADY = 3, BAR_ACC = 5;

__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

}  // namespace

// wpk: bf16 [mat: Whh_hi, Whh_lo, Wih_hi, Wih_lo][tile 2][lane 128][64]; bias_p: fp32 [tile 2][lane 128]
// (both pre-scaled by -log2 e / -2 log2 e, see launch_pack_lstm_tc); xmap: the planar input as a 5-d tensor
template <bool PROF>
__global__ void __launch_bounds__(THREADS, 1)
    lstm_tc_kernel(const __grid_constant__ CUtensorMap xmap, const uint4 *__restrict__ wpk,
                   const float *__restrict__ bias_p, const float *__restrict__ chan_add, uint4 *__restrict__ y, int B,
                   int T, int opts, long long *__restrict__ prof) {
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t s_base = smem_u32(smem);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + OFF_BAR + 2 * 64);
  // A/B switch (wm_debug_lstm_opts bit 0): the epilogue warps wait on the accumulator's mbarrier themselves (a
  // try_wait with a suspend hint sleeps until the tcgen05.commit arrives) instead of being released by the MMA thread
  // through a named barrier after ITS wait
  const bool direct_acc = (opts & 1) != 0;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const size_t RP = (size_t)T + 2 * PAD;
  const int nchunk = (T + TC_STEPS - 1) / TC_STEPS;

  if (tid == 0) {
    for (int g = 0; g < NGRP; ++g) {
      const uint32_t bars = s_base + OFF_BAR + 64 * g;
      for (int i = 0; i < 2; ++i) {
        mbar_init(bars + 8 * i, 1);                      // acc_full
        mbar_init(bars + 16 + 8 * i, 1);                 // x_full
        mbar_init(bars + 32 + 8 * i, 1);                 // x_empty
      }
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // h_0 = 0
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
  const int role = warp < W_MMA0 ? 0 : 1;          // 0 epilogue, 1 MMA issuer (+ TMA producer)
  const int g = role == 0 ? warp / (N_EPI / 32) : warp - W_MMA0;
  const int b0 = blockIdx.x * (NGRP * NCL) + g * NCL;      // first clip of the group
  const int nb = max(0, min(NCL, B - b0));
  const uint32_t gs = s_base + g * G_BYTES;
  uint8_t *gsm = smem + g * G_BYTES;
  const uint32_t bars = s_base + OFF_BAR + 64 * g;
  const uint32_t acc_full0 = bars, x_full0 = bars + 16, x_empty0 = bars + 32;
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
      if (direct_acc) mbar_wait(acc_full0 + 8 * buf, (uint32_t)((t >> 1) & 1));
      else bar_sync(BAR_ACC + g, N_EPI + 32);               // the MMA thread saw this step's accumulator complete
      tc_fence_after();
      long long c1 = pf ? clock64() : 0;
      uint32_t r[32];
      tmem_ld32(tm_acc + buf * 64 + m * 32 + ((uint32_t)(q * 32) << 16), r);
      tmem_ld_wait();
      tc_fence_before();
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
      bar_sync(BAR_PHASE + g, N_EPI);
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
      // hand h_t to the MMA warp through a hardware named barrier (arrive here, sync there): the barrier orders
      // these stores before the MMA thread's proxy fence
      bar_arrive(BAR_HREADY + g, N_EPI + 32);
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
    // ===================== MMA issuer + TMA producer (one elected lane works; the warp only relays barriers) =====
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
    auto load_x = [&](int ch) {   // steps 8 ch .. 8 ch + 7 of the group's 16 clips -> stage ch & 1
      const int st = ch & 1;
      mbar_arrive_expect_tx(x_full0 + 8 * st, XSTAGE);
      tma_load_5d(gs + G_X + st * XSTAGE, &xmap, x_full0 + 8 * st, 0, b0, 0, 0, PAD + ch * TC_STEPS);
    };
    const bool pf = PROF && prof != nullptr && blockIdx.x == 0 && g == 0 && issuer;
    long long pm[4] = {0, 0, 0, 0};
    // (waits and MMA issue live in separate single-lane regions: a region that mixes a polling loop with the MMAs
    // makes ptxas wrap every tcgen05.mma in its own operand-uniformising loop)
    if (issuer) {
      load_x(0);
      if (nchunk > 1) load_x(1);
      mbar_wait(x_full0, 0);
    }
    __syncwarp();
    tc_fence_after();
    if (issuer) issue(smem_desc(gs + G_X, 2 * NCL * 16, 128), 2, 0, 0);     // x part of step 0
    __syncwarp();
#pragma unroll 1
    for (int t = 0; t < T; ++t) {
      const int buf = t & 1;
      long long m0 = pf ? clock64() : 0;
      if (t > 0) {
        bar_sync(BAR_HREADY + g, N_EPI + 32);               // h_{t-1} is in shared memory
        if (issuer) fence_async_smem();                     // the epilogue warps' stores -> the MMA's operand fetch
        tc_fence_after();
      }
      long long m1 = pf ? clock64() : 0;
      if (issuer) {
        issue(h_desc, 0, buf, 1);                           // + W_hh . h_{t-1}
        tc_commit(acc_full0 + 8 * buf);
      }
      __syncwarp();
      long long m2 = pf ? clock64() : 0;
      if (!direct_acc) {
        if (issuer) mbar_wait(acc_full0 + 8 * buf, (uint32_t)((t >> 1) & 1));
        __syncwarp();
        tc_fence_before();
        bar_arrive(BAR_ACC + g, N_EPI + 32);                // release the epilogue warps
      }
      long long m3 = pf ? clock64() : 0;
      const int t1 = t + 1;
      if (t1 < T) {
        // W_ih . x_{t+1} into the other accumulator (drained by the epilogue before it handed over h_{t-1})
        const int ch = t1 / TC_STEPS, tt = t1 % TC_STEPS, st = ch & 1;
        if (tt == 0) {
          if (issuer) mbar_wait(x_full0 + 8 * st, (uint32_t)((ch >> 1) & 1));
          __syncwarp();
        }
        tc_fence_after();
        if (issuer) {
          issue(smem_desc(gs + G_X + st * XSTAGE + tt * XTILE, 2 * NCL * 16, 128), 2, buf ^ 1, 0);
          if (tt == TC_STEPS - 1) tc_commit(x_empty0 + 8 * st);
        }
        __syncwarp();
        if (tt == 1 && ch >= 1 && ch + 1 < nchunk) {        // chunk ch - 1's stage has been read: refill it with chunk ch + 1
          if (issuer) {
            mbar_wait(x_empty0 + 8 * (st ^ 1), (uint32_t)(((ch - 1) >> 1) & 1));
            load_x(ch + 1);
          }
          __syncwarp();
        }
      }
      if (pf) { long long m4 = clock64(); pm[0] += m1 - m0; pm[1] += m2 - m1; pm[2] += m3 - m2; pm[3] += m4 - m3; }
    }
    if (pf) {
      for (int i = 0; i < 4; ++i) prof[8 + i] = pm[i];
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
constexpr int kLstmDefaultOpts = 0;
static int g_lstm_opts = kLstmDefaultOpts;
void set_lstm_opts(int o) { g_lstm_opts = o < 0 ? kLstmDefaultOpts : o; }
int get_debug_opts() { return g_lstm_opts; }
void set_lstm_profile_buffer(long long *p) { g_lstm_prof = p; }
long long *get_profile_buffer() { return g_lstm_prof; }

// The planar input [clip][plane = hi/lo x chunk][row T + 2 PAD][16 B] as a 5-d tensor of 32-bit words with the
// dimensions ordered the way a step's B tile wants them in shared memory: {4 words, clip, hi/lo, chunk, row}.
static int make_x_map(CUtensorMap *map, const void *x, int B, int T) {
  static PFN_cuTensorMapEncodeTiled_v12000 encode = nullptr;
  if (encode == nullptr) {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    WM_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (fn == nullptr || qres != cudaDriverEntryPointSuccess) {
      set_error("cuTensorMapEncodeTiled is not available in this driver");
      return -3;
    }
    encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  }
  const cuuint64_t RP = (cuuint64_t)T + 2 * PAD;
  const cuuint64_t dims[5] = {4, (cuuint64_t)B, 2, 8, RP};
  const cuuint64_t strides[4] = {16 * RP * 16, 8 * RP * 16, RP * 16, 16};   // bytes, dimensions 1..4
  const cuuint32_t box[5] = {4, NCL, 2, 8, TC_STEPS};
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 5, const_cast<void *>(x), dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) for the LSTM input (B=%d, T=%d)", (int)r, B, T);
    return -3;
  }
  return 0;
}

int launch_lstm_tc(const void *x, const void *wpk, const float *bias_p, const float *chan_add, void *y, int B, int T,
                   cudaStream_t st) {
  if (B == 0 || T == 0) return 0;
  static bool attr_set = false;
  if (!attr_set) {
    WM_CHECK_CUDA(cudaFuncSetAttribute(lstm_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, LSTM_SMEM));
    WM_CHECK_CUDA(cudaFuncSetAttribute(lstm_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, LSTM_SMEM));
    attr_set = true;
  }
  CUtensorMap xmap;
  WM_TRY(make_x_map(&xmap, x, B, T));
  const int grid = (B + NGRP * NCL - 1) / (NGRP * NCL);
  if (g_lstm_prof != nullptr)   // developer build of the same kernel with per-phase cycle counters (tools/lstm_profile.py)
    lstm_tc_kernel<true><<<grid, THREADS, LSTM_SMEM, st>>>(xmap, reinterpret_cast<const uint4 *>(wpk), bias_p, chan_add,
                                                           reinterpret_cast<uint4 *>(y), B, T, g_lstm_opts, g_lstm_prof);
  else
    lstm_tc_kernel<false><<<grid, THREADS, LSTM_SMEM, st>>>(xmap, reinterpret_cast<const uint4 *>(wpk), bias_p, chan_add,
                                                            reinterpret_cast<uint4 *>(y), B, T, g_lstm_opts, nullptr);
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
