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
// Warps: 0..7 group 1 (conv1 accumulator -> intermediate tile), 8..15 group 2 (conv2 accumulator
// + residual -> y); in both, TMEM lane quadrant = warp % 4 and 32-channel half = (warp / 4) % 2.
// Warp 16 producer (bulk copies), warp 17 MMA issuer + TMEM owner.  TMEM: conv1 and
// conv2 accumulators double buffered (4 x 128 columns).  Shared memory: both weight images
// (96 KB), 2 x-tile stages, 2 intermediate tiles.  The MMA warp issues conv1 of tile i+1 before
// conv2 of tile i, so the tensor pipe works on the next tile while the epilogue warps turn tile
// i's conv1 accumulator into conv2's operand.
#include <cuda_bf16.h>
#include <stdlib.h>
#include <string.h>

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
constexpr int NST = 2;                     // x-tile stages (a stage is free as soon as conv1 has read it)
constexpr int NU = 2;                      // intermediate tiles (conv2 of tile i overlaps the refill for tile i+1)
constexpr int OFF_W = 0;
constexpr int OFF_X = OFF_W + 2 * W_IMG_B;
constexpr int OFF_U = OFF_X + NST * TILE_B;
constexpr int OFF_BIAS = OFF_U + NU * TILE_B;
constexpr int OFF_BAR = OFF_BIAS + 512;
constexpr int RB_SMEM = OFF_BAR + 256;
constexpr uint32_t kIdesc = make_idesc(128, 128);
// The lo x lo partial product is ~2^-18 of the result, far below the parity budget: the A_lo MMA multiplies
// only the first 64 columns of B (= W_hi), i.e. 3 products in 1.5 MMAs' worth of tensor time.
#ifndef WM_FULL_PRODUCTS
constexpr uint32_t kIdescLo = make_idesc(128, 64);
#else
constexpr uint32_t kIdescLo = kIdesc;
#endif
// CTA-pair mode: per (tap, k chunk) a CTA holds its half of [W_hi | W_lo] (64 rows, operand of the A_hi MMA) and its
// half of W_hi (32 rows, operand of the A_lo MMA)
constexpr int W2_CHUNK_B = (64 + 32) * 16;             // 1 536
constexpr int W2_IMG_B = 3 * 8 * W2_CHUNK_B;           // 36 864 per convolution
constexpr uint32_t kIdesc2 = make_idesc(256, 128), kIdescLo2 = make_idesc(256, 64);
constexpr int N_GRP = 256;                  // threads per epilogue group (8 warps: 4 lane quadrants x 2 channel halves)
constexpr int W_PROD = 2 * N_GRP / 32, W_MMA = W_PROD + 1, RB_THREADS = 2 * N_GRP + 64;
static_assert(RB_SMEM <= 232448, "shared memory budget");

}  // namespace

// Biases and 1x1 head weights travel BY VALUE as a kernel parameter: parameters live in the constant bank, so every
// bias add / head FFMA takes its operand as a constant (no load instruction, no latency to hide, no state shared
// between launches).  Measured: with the head weights fetched through L1 the 17-output epilogue stalled on every
// FFMA and the kernel ran 2.4x slower than the plain ResBlock; with the biases in shared memory their (scalar,
// predicated) loads were 60 % of the kernel's LSU shared-memory wavefronts (profiles/r2_resblock_smem_wavefronts.md).
template <int NW>
struct RbParams {
  float b1[64], b2[64];   // folded BatchNorm biases of conv1 / conv2 (used when BIASP)
  float w[NW * 64];       // 1x1 head: row 0 as [64]; rows 1.. (NW > 1, the vote epilogue) transposed to [channel][NW - 1]
  float b[NW];
};
// The message-head FMAs of the fused vote epilogue for the 16-channel slice P.  Compile-time weight indices: the
// weights are uniform-register operands fetched by LDCU.128 (message rows are stored [channel][bit] in the kernel
// parameter, so a channel's 16 weights are one 64-byte run), and two bits share one packed FFMA2 (fma.rn.f32x2: the
// activation is the broadcast scalar operand, the weight pair a uniform-register pair) — 512 FFMA2 + 256 LDCU.128 per
// row.  With a run-time slice index the same FMAs needed an indexed LDC.64 per two weights and this epilogue, not the
// tensor pipe, paced the kernel.  Detector pass with votes, 4096 clips: 28 ms -> 19.0 (scalar FFMA, WM_VOTE_F32X2=0)
// -> 18.0 ms; 15.0 ms without votes.  Each lane of the packed FMA is an ordinary fp32 FMA and per bit the channels are
// accumulated in ascending order, so the vote counts are those of the scalar form.
#ifndef WM_VOTE_F32X2
#define WM_VOTE_F32X2 1
#endif
constexpr int NVOTE = WM_FUSED_VOTE_MAX - 1;
template <int P>
__device__ __forceinline__ void vote_fma(const float (&o)[16], float (&lacc)[NVOTE], const float *w) {
#if WM_VOTE_F32X2
  unsigned long long acc2[NVOTE / 2];
#pragma unroll
  for (int jp = 0; jp < NVOTE / 2; ++jp) asm("mov.b64 %0, {%1, %2};" : "=l"(acc2[jp]) : "f"(lacc[2 * jp]), "f"(lacc[2 * jp + 1]));
#pragma unroll
  for (int c = 0; c < 16; ++c) {
    unsigned long long ov;
    asm("mov.b64 %0, {%1, %1};" : "=l"(ov) : "f"(o[c]));
#pragma unroll
    for (int jp = 0; jp < NVOTE / 2; ++jp) {
      const unsigned long long wv = *reinterpret_cast<const unsigned long long *>(&w[64 + (P * 16 + c) * NVOTE + 2 * jp]);
      asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc2[jp]) : "l"(ov), "l"(wv));
    }
  }
#pragma unroll
  for (int jp = 0; jp < NVOTE / 2; ++jp) asm("mov.b64 {%0, %1}, %2;" : "=f"(lacc[2 * jp]), "=f"(lacc[2 * jp + 1]) : "l"(acc2[jp]));
#else
#pragma unroll
  for (int c = 0; c < 16; ++c) {
#pragma unroll
    for (int j = 0; j < NVOTE; ++j) lacc[j] = fmaf(o[c], w[64 + (P * 16 + c) * NVOTE + j], lacc[j]);
  }
#endif
}

template <int NHEAD>
struct RbParamsFor { using type = RbParams<NHEAD == 3 ? WM_FUSED_VOTE_MAX : 1>; };

// NHEAD = 0: plain ResBlock.  NHEAD = 1: + Conv1d(64,1,1) -> head_out[b][t] (py/main16.py:146).
// NHEAD = 3: NHEAD = 2 plus the per-sample message logits (all rows of the head, constant-bank FFMAs) reduced to
// per-(tile, warp) counts of positive logits: the majority vote of evaluate_model (py/main16.py:398) without a
// (B,T,1+bits) logits tensor.
// BIASP: biases come from the kernel parameter (callers that hold a host copy), else from shared memory.
// NHEAD = 2 (detector): + channel 0 of Conv1d(64,1+bits,1) -> head_out[b][t] = sigmoid(logit 0), and per
// (tile, warp) partial sums over the valid samples of that probability and of the 64 ACTIVATIONS: the message
// logits are only ever used as means over time (py/main16.py:1142-1146), and a mean of a linear map is the
// linear map of the mean, so the 16 x 64 message head is applied once per clip by detect_finalize_kernel
// instead of 1024 FMAs per sample here.
// CTA2: the kernel runs as clusters of two CTAs (one TPC); each CTA owns its own tile (128 rows of M) and all its
// epilogue work, but the MMAs are issued by the even CTA for both (tcgen05 cta_group::2, M = 256) with every
// weight operand split between the two shared memories: 21 % fewer operand bytes per CTA, which is what this
// kernel is bound by.  The odd CTA's MMA warp relays its "tile landed" barriers to the even CTA; its epilogue
// warps arrive on the even CTA's barriers directly; MMA completions are multicast to both.
template <int NHEAD, bool CTA2, bool BIASP>
// 18 warps = 5 on one SM sub-partition (16 K registers each): 96 registers per thread is the ceiling
__global__ void __launch_bounds__(RB_THREADS, 1)
    resblock_tc_kernel(const uint4 *__restrict__ x, const uint4 *__restrict__ w_img, const float *__restrict__ b1,
                       const float *__restrict__ b2, uint4 *__restrict__ y, float *__restrict__ y32, int B, int T,
                       const __grid_constant__ typename RbParamsFor<NHEAD>::type hp, int nvote,
                       float *__restrict__ head_out, float *__restrict__ partials, const int *__restrict__ valid_len,
                       long long *__restrict__ prof) {
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t s_base = smem_u32(smem);
  const uint32_t w_smem = s_base + OFF_W, x_smem = s_base + OFF_X, u_smem = s_base + OFF_U;
  const uint32_t bars = s_base + OFF_BAR;
  // barriers (8 B each), tmem slot behind them
  enum { FULL = 0, EMPTY = 2, WBAR = 4, D1_FULL = 5, D2_FULL = 7, D2_EMPTY = 9, U_FULL = 11, U_EMPTY = 13,
         // CTA-pair mode, even CTA only: the odd CTA's FULL / WBAR / U_FULL / D2_EMPTY events, forwarded by its relay thread
         PFULL = 15, PWBAR = 17, PU_FULL = 18, PD2_EMPTY = 20, NBAR = 22 };
  auto bar = [&](int i) { return bars + 8 * i; };
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + OFF_BAR + 8 * NBAR);
  float *bias_s = reinterpret_cast<float *>(smem + OFF_BIAS);   // [0..63] b1, [64..127] b2

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntile_t = (T + TO - 1) / TO;
  const long long ntiles = (long long)B * ntile_t;
  const size_t RP = (size_t)T + 2 * PAD;
  // tile of iteration i (clamped; a pair's second tile may not exist) and the number of iterations
  const uint32_t rank = CTA2 ? cluster_ctarank() : 0;
  const long long stride_t = CTA2 ? (long long)(gridDim.x >> 1) : (long long)gridDim.x;
  const long long first_t = CTA2 ? (long long)(blockIdx.x >> 1) : (long long)blockIdx.x;
  const long long nunits = CTA2 ? (ntiles + 1) / 2 : ntiles;
  const long long my_tiles = nunits > first_t ? (nunits - first_t + stride_t - 1) / stride_t : 0;
  auto tile_raw = [&](long long i) { return CTA2 ? 2 * (first_t + i * stride_t) + rank : first_t + i * stride_t; };
  auto tile_of = [&](long long i) { const long long t_ = tile_raw(i); return t_ < ntiles ? t_ : ntiles - 1; };

  if (threadIdx.x == 0) {
    for (int s = 0; s < NST; ++s) { mbar_init(bar(FULL + s), 1); mbar_init(bar(EMPTY + s), 1); }
    mbar_init(bar(WBAR), 1);
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar(D1_FULL + a), 1); mbar_init(bar(D2_FULL + a), 1); mbar_init(bar(D2_EMPTY + a), 4);
      mbar_init(bar(U_FULL + a), N_GRP / 32); mbar_init(bar(U_EMPTY + a), 1);
      mbar_init(bar(PFULL + a), 1); mbar_init(bar(PU_FULL + a), 1); mbar_init(bar(PD2_EMPTY + a), 1);
    }
    mbar_init(bar(PWBAR), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (!BIASP && threadIdx.x < 128) bias_s[threadIdx.x] = threadIdx.x < 64 ? b1[threadIdx.x] : b2[threadIdx.x - 64];
  if (warp == W_MMA) {
    if constexpr (CTA2) {   // the same warp of both CTAs, same destination offset
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CTA2) cluster_sync_all();   // both CTAs' barriers exist before anyone arrives on or multicasts to them
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == W_PROD) {
    // ===== producer: weights once, then one x tile (16 planes x 130 rows) per stage =====
    if (lane == 0) {
      if constexpr (CTA2) {
        mbar_arrive_expect_tx(bar(WBAR), 2 * W2_IMG_B);
        const uint8_t *img = reinterpret_cast<const uint8_t *>(w_img);
        for (int cj = 0; cj < 2 * 3 * 8; ++cj) {   // (conv, tap, k chunk): image chunk = 128 rows x 16 B
          bulk_g2s(w_smem + cj * W2_CHUNK_B, img + ((size_t)cj * 128 + 64 * rank) * 16, 64 * 16, bar(WBAR));
          bulk_g2s(w_smem + cj * W2_CHUNK_B + 64 * 16, img + ((size_t)cj * 128 + 32 * rank) * 16, 32 * 16, bar(WBAR));
        }
      } else {
        mbar_arrive_expect_tx(bar(WBAR), 2 * W_IMG_B);
        for (int j = 0; j < 6; ++j)
          bulk_g2s(w_smem + j * W_TAP_B, reinterpret_cast<const uint8_t *>(w_img) + (size_t)j * W_TAP_B, W_TAP_B, bar(WBAR));
      }
      TileWalk wk(first_t, stride_t, ntile_t);       // single-CTA mode: tile_raw(i) < ntiles for every i < my_tiles
      for (long long i = 0; i < my_tiles; ++i, wk.next()) {
        const int s = (int)(i % NST);
        mbar_wait(bar(EMPTY + s), (uint32_t)(((i / NST) & 1) ^ 1));
        mbar_arrive_expect_tx(bar(FULL + s), TILE_B);
        long long b = wk.b;
        int t0 = wk.tt * TO;
        if constexpr (CTA2) {
          const long long tile = tile_of(i);
          b = tile / ntile_t;
          t0 = (int)(tile % ntile_t) * TO;
        }
        const size_t row0 = (size_t)(t0 + PAD - 2);
#pragma unroll 4
        for (int p = 0; p < 16; ++p)
          bulk_g2s(x_smem + s * TILE_B + p * PLANE_B, x + ((size_t)(b * 16 + p) * RP + row0), PLANE_B, bar(FULL + s));
      }
    }
    __syncwarp();
  } else if (warp == W_MMA && CTA2 && rank == 1) {
    // ===== odd CTA of a pair: no MMA issue.  One thread forwards this CTA's "operand ready / accumulator drained"
    // events to the even CTA, in the order its MMA thread waits for them.  The epilogue warps arrive on their own
    // CTA's barriers as in single-CTA mode: a cluster-scope release from a warp with global stores in flight costs
    // a full store drain (measured: +70 % epilogue time), from this idle thread it costs nothing.
    if (elect_one()) {
      mbar_wait(bar(WBAR), 0);
      mbar_arrive_cluster(bar(PWBAR), 0);
      if (my_tiles > 0) {
        mbar_wait(bar(FULL), 0);
        mbar_arrive_cluster(bar(PFULL), 0);
      }
      for (long long i = 0; i < my_tiles; ++i) {
        const int a = (int)(i & 1);
        if (i + 1 < my_tiles) {
          const long long n = i + 1;
          const int s = (int)(n % NST);
          mbar_wait(bar(FULL + s), (uint32_t)((n / NST) & 1));
          mbar_arrive_cluster(bar(PFULL + s), 0);
        }
        mbar_wait(bar(U_FULL + a), (uint32_t)((i >> 1) & 1));
        mbar_arrive_cluster(bar(PU_FULL + a), 0);
        if (i >= 2) {
          mbar_wait(bar(D2_EMPTY + a), (uint32_t)(((i >> 1) - 1) & 1));
          mbar_arrive_cluster(bar(PD2_EMPTY + a), 0);
        }
      }
    }
    __syncwarp();
  } else if (warp == W_MMA) {
    // ===== MMA issuer: one elected lane waits, issues and commits; the other lanes idle =====
    if (elect_one()) {
      auto conv = [&](uint32_t a_tile, uint32_t w_img_s, uint32_t d_tmem) {
        const uint64_t a0 = smem_desc(a_tile, PLANE_B, 128);
        const uint64_t b0 = CTA2 ? smem_desc(w_img_s, W2_CHUNK_B, 128) : smem_desc(w_img_s, 2048, 128);
#pragma unroll
        for (int j = 0; j < 3; ++j) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              const uint64_t ad = a0 + (uint64_t)(((half * 8 + 2 * kk) * PLANE_B + j * 16) >> 4);
              const uint32_t acc = (j | kk | half) != 0 ? 1u : 0u;
              if constexpr (CTA2)
                mma_bf16_cta2(d_tmem, ad, b0 + (uint64_t)(((j * 8 + 2 * kk) * W2_CHUNK_B + half * 64 * 16) >> 4),
                              half ? kIdescLo2 : kIdesc2, acc);
              else
                mma_bf16(d_tmem, ad, b0 + (uint64_t)((j * W_TAP_B + (2 * kk) * 2048) >> 4), half ? kIdescLo : kIdesc, acc);
            }
          }
        }
      };
      auto commit = [&](uint32_t b_) {
        if constexpr (CTA2) tc_commit_cta2(b_);
        else tc_commit(b_);
      };
      auto wait_x = [&](int s, uint32_t ph) {   // the x tile of stage s has landed (in both CTAs)
        mbar_wait(bar(FULL + s), ph);
        if constexpr (CTA2) mbar_wait_cluster(bar(PFULL + s), ph);
      };
      auto wait_both = [&](int local, int peer, uint32_t ph) {   // an event of this CTA's epilogue warps and of the peer's
        mbar_wait(bar(local), ph);
        if constexpr (CTA2) mbar_wait_cluster(bar(peer), ph);
      };
      const bool pf = prof != nullptr && blockIdx.x == 0;
      long long pm[5] = {0, 0, 0, 0, 0};
      mbar_wait(bar(WBAR), 0);
      if constexpr (CTA2) mbar_wait_cluster(bar(PWBAR), 0);
      constexpr int W_CONV2 = CTA2 ? W2_IMG_B : W_IMG_B;
      if (my_tiles > 0) {
        wait_x(0, 0);
        tc_fence_after();
        conv(x_smem, w_smem, tmem);
        commit(bar(D1_FULL));
        commit(bar(EMPTY));      // conv1 is the x stage's only reader: hand it back to the producer
      }
      for (long long i = 0; i < my_tiles; ++i) {
        const int a = (int)(i & 1);
        long long m0 = pf ? clock64() : 0, m1 = m0, m2 = m0;
        if (i + 1 < my_tiles) {  // conv1 of the next tile; D1[an] was drained by E1(i-1), implied by u_full(i-1)
          const long long n = i + 1;
          const int s = (int)(n % NST), an = (int)(n & 1);
          wait_x(s, (uint32_t)((n / NST) & 1));
          m1 = pf ? clock64() : 0;
          tc_fence_after();
          conv(x_smem + s * TILE_B, w_smem, tmem + an * 128);
          commit(bar(D1_FULL + an));
          commit(bar(EMPTY + s));
          m2 = pf ? clock64() : 0;
        }
        // conv2 of this tile
        wait_both(U_FULL + a, PU_FULL + a, (uint32_t)((i >> 1) & 1));
        if (i >= 2) wait_both(D2_EMPTY + a, PD2_EMPTY + a, (uint32_t)(((i >> 1) - 1) & 1));
        long long m3 = pf ? clock64() : 0;
        tc_fence_after();
        conv(u_smem + a * TILE_B, w_smem + W_CONV2, tmem + 256 + a * 128);
        commit(bar(D2_FULL + a));
        commit(bar(U_EMPTY + a));
        if (pf) { long long m4 = clock64(); pm[0] += m1 - m0; pm[1] += m2 - m1; pm[2] += m3 - m2; pm[3] += m4 - m3; }
      }
      if (pf) {
        for (int k = 0; k < 4; ++k) prof[16 + k] = pm[k];
        prof[31] = my_tiles;
      }
    }
    __syncwarp();
  } else if (warp < N_GRP / 32) {
    // ===== group 1: conv1 accumulator -> intermediate tile (conv2's A operand) =====
    const int q = warp & 3;                            // TMEM lane quadrant; 32-channel half = warp >> 2
    const int row = q * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const bool pfe = prof != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
    long long pe[3] = {0, 0, 0};
    // ONE copy of this loop for both 32-channel halves (runtime `half`): duplicating it per half (to make the bias
    // indices immediates) pushed the kernel's hot code out of the instruction cache -- 48 % of all warp stalls became
    // "no instruction" and the kernel ran 37 % longer (profiles/r2_resblock_icache.md).  With BIASP the biases are
    // indexed constant-bank loads (LDC), which do not touch the shared-memory pipeline the MMAs are bound by.
    const int half = warp >> 2;
    TileWalk wk(first_t, stride_t, ntile_t);
    for (long long i = 0; i < my_tiles; ++i, wk.next()) {
      int t0 = wk.tt * TO;
      if constexpr (CTA2) t0 = (int)(tile_of(i) % ntile_t) * TO;
      const int a = (int)(i & 1);
      const int tu = t0 - 1 + row;
      const bool inside = tu >= 0 && tu < T;     // conv2 zero-pads the intermediate feature map
      long long e0 = pfe ? clock64() : 0;
      uint8_t *us = smem + OFF_U + a * TILE_B + row * 16;
      mbar_wait_warp(bar(D1_FULL + a), (uint32_t)((i >> 1) & 1));
      long long e1 = pfe ? clock64() : 0;
      if (i >= 2) mbar_wait_warp(bar(U_EMPTY + a), (uint32_t)(((i >> 1) - 1) & 1));   // conv2(i-2) has finished reading U[a]
      long long e2 = pfe ? clock64() : 0;
      tc_fence_after();
      const uint32_t taddr = tmem + a * 128 + lane_off;
#pragma unroll 1
      for (int pp = 0; pp < 2; ++pp) {
        const int p = half * 2 + pp;
        float v1[16], v2[16];
        tmem_ld16(taddr + p * 16, v1);
        tmem_ld16(taddr + 64 + p * 16, v2);
        tmem_ld_wait();
        float o[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          const float bv = BIASP ? hp.b1[p * 16 + c] : bias_s[p * 16 + c];
          const float v = fmaxf(v1[c] + v2[c] + bv, 0.0f);
          o[c] = inside ? v : 0.0f;
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int ch = p * 2 + h;
          uint4 hi, lo;
          split8(o + h * 8, hi, lo);
          *reinterpret_cast<uint4 *>(us + ch * PLANE_B) = hi;
          *reinterpret_cast<uint4 *>(us + (8 + ch) * PLANE_B) = lo;
        }
      }
      tc_fence_before();
      fence_async_smem();
      mbar_arrive_warp(bar(U_FULL + a));
      if (pfe) { long long e3 = clock64(); pe[0] += e1 - e0; pe[1] += e2 - e1; pe[2] += e3 - e2; }
    }
    if (pfe) {
      for (int k = 0; k < 3; ++k) prof[24 + k] = pe[k];
    }
  } else {
    // ===== group 2: conv2 accumulator + bias + residual -> ReLU -> y (and the fused 1x1 head) =====
    // two sub-groups of 4 warps (one per TMEM lane quadrant) take alternate tiles, so a thread owns a whole
    // row (64 channels, four 16-channel passes) and has two tile periods for it; D2[g] belongs to sub-group g
    const int w2 = warp - N_GRP / 32;
    const int q = w2 & 3, g = w2 >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const bool pfe = prof != nullptr && blockIdx.x == 0 && lane == 0 && w2 == 0;
    long long pe[3] = {0, 0, 0};
    TileWalk wk(first_t + g * stride_t, 2 * stride_t, ntile_t);
    for (long long i = g; i < my_tiles; i += 2, wk.next()) {
      bool real_tile = true;
      long long b = wk.b;
      int tt = wk.tt;
      if constexpr (CTA2) {
        const long long tile = tile_of(i);
        real_tile = tile_raw(i) < ntiles;               // a pair's second tile may not exist: compute, do not store
        b = tile / ntile_t;
        tt = (int)(tile % ntile_t);
      }
      const int t0 = tt * TO;
      const int t = t0 + row;
      const bool live = real_tile && row < TO && t < T;
      const size_t prow = (size_t)t + PAD;
      long long e0 = pfe ? clock64() : 0;
      if (y != nullptr && real_tile && q == 0 && lane < 2 * PAD) {  // the planes' zero padding rows
        const bool head = lane < PAD;
        if (head ? (t0 == 0) : (t0 + TO >= T)) {
          const size_t zr = head ? (size_t)lane : (size_t)T + lane;
          for (int pl = 0; pl < 16; ++pl) y[((size_t)(b * 16 + pl)) * RP + zr] = make_uint4(0, 0, 0, 0);
        }
      }
      // residual x[t] comes from L2 (the tile was fetched a moment ago), 16 channels (4 x 16 B) per pass,
      // requested one pass ahead; the first request goes out before the accumulator wait
      uint4 rres[4];
      auto fetch = [&](int p) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int ch = p * 2 + h;
          rres[2 * h] = live ? __ldg(&x[((size_t)(b * 16 + ch)) * RP + prow]) : make_uint4(0, 0, 0, 0);
          rres[2 * h + 1] = live ? __ldg(&x[((size_t)(b * 16 + 8 + ch)) * RP + prow]) : make_uint4(0, 0, 0, 0);
        }
      };
      fetch(0);
      mbar_wait_warp(bar(D2_FULL + g), (uint32_t)((i >> 1) & 1));
      long long e1 = pfe ? clock64() : 0;
      tc_fence_after();
      const uint32_t taddr = tmem + 256 + g * 128 + lane_off;
      float hacc = 0.0f;
      [[maybe_unused]] bool counted = false;
      [[maybe_unused]] float *pdst = nullptr;
      [[maybe_unused]] float lacc[NHEAD == 3 ? WM_FUSED_VOTE_MAX - 1 : 1];
      if constexpr (NHEAD == 3) {
#pragma unroll
        for (int j = 0; j < WM_FUSED_VOTE_MAX - 1; ++j) lacc[j] = hp.b[1 + j];
      }
      if constexpr (NHEAD >= 2) {
        const int vl = valid_len != nullptr ? min(max(valid_len[b], 0), T) : T;
        counted = live && t < vl;
        pdst = partials + (((size_t)b * ntile_t + tt) * 4 + q) * WM_DET_PART;
      }
      // NOT unrolled: four copies of this body (~5 KB each) pushed the kernel's hot code (MMA issue + both epilogue
      // groups + producer) past the 32 KB instruction cache; bias / head operands become indexed constant loads
#pragma unroll 1
      for (int p = 0; p < 4; ++p) {
        float v1[16], v2[16], o[16];
        tmem_ld16(taddr + p * 16, v1);
        tmem_ld16(taddr + 64 + p * 16, v2);
        tmem_ld_wait();
        if (p == 3) {
          tc_fence_before();
          mbar_arrive_warp(bar(D2_EMPTY + g));     // D2[g] may be overwritten by conv2(i+2)
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float r[8];
          join8(rres[2 * h], rres[2 * h + 1], r);
#pragma unroll
          for (int c = 0; c < 8; ++c)
            o[h * 8 + c] = fmaxf(v1[h * 8 + c] + v2[h * 8 + c] +
                                     (BIASP ? hp.b2[p * 16 + h * 8 + c] : bias_s[64 + p * 16 + h * 8 + c]) + r[c], 0.0f);
        }
        if (p < 3) fetch(p + 1);
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
        if constexpr (NHEAD > 0) {
          // one output of the 1x1 head: weights are constant-bank operands of the FFMAs (kernel parameter)
#pragma unroll
          for (int c = 0; c < 16; ++c) hacc = fmaf(o[c], hp.w[p * 16 + c], hacc);
        }
        if constexpr (NHEAD == 3) {
          // per-sample message logits (rows 1.. of the head): one specialisation per channel slice, see vote_fma
          switch (p) {
            case 0: vote_fma<0>(o, lacc, hp.w); break;
            case 1: vote_fma<1>(o, lacc, hp.w); break;
            case 2: vote_fma<2>(o, lacc, hp.w); break;
            default: vote_fma<3>(o, lacc, hp.w); break;
          }
        }
        if constexpr (NHEAD >= 2) {
          // sum of these 16 channels over the warp's 32 rows by recursive halving: after the steps 16, 8, 4, 2
          // lane l holds channel l / 2 summed over 16 lanes, the last exchange folds the lane pair
          float a8[8], a4[4], a2[2];
          const bool b16 = lane & 16, b8 = lane & 8, b4 = lane & 4, b2 = lane & 2;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float lo_v = counted ? o[k] : 0.0f, hi_v = counted ? o[8 + k] : 0.0f;
            a8[k] = (b16 ? hi_v : lo_v) + __shfl_xor_sync(0xffffffffu, b16 ? lo_v : hi_v, 16);
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) a4[k] = (b8 ? a8[4 + k] : a8[k]) + __shfl_xor_sync(0xffffffffu, b8 ? a8[k] : a8[4 + k], 8);
#pragma unroll
          for (int k = 0; k < 2; ++k) a2[k] = (b4 ? a4[2 + k] : a4[k]) + __shfl_xor_sync(0xffffffffu, b4 ? a4[k] : a4[2 + k], 4);
          float a1 = (b2 ? a2[1] : a2[0]) + __shfl_xor_sync(0xffffffffu, b2 ? a2[0] : a2[1], 2);
          a1 += __shfl_xor_sync(0xffffffffu, a1, 1);
          if (real_tile && (lane & 1) == 0) pdst[p * 16 + (lane >> 1)] = a1;
        }
      }
      if constexpr (NHEAD == 1) {
        if (live) head_out[(size_t)b * T + t] = hacc + hp.b[0];
      } else if constexpr (NHEAD >= 2) {
        if constexpr (NHEAD == 3) {   // positive message logits among this warp's counted samples, per bit
#pragma unroll
          for (int j = 0; j < WM_FUSED_VOTE_MAX - 1; ++j) {
            const unsigned m = __ballot_sync(0xffffffffu, counted && j < nvote && lacc[j] > 0.0f);
            if (real_tile && lane == 0) pdst[WM_DET_VOTE0 + j] = (float)__popc(m);
          }
        }
        const float pr = sigmoid_acc(hacc + hp.b[0]);
        if (live && head_out != nullptr) head_out[(size_t)b * T + t] = pr;
        float red = counted ? pr : 0.0f;
#pragma unroll
        for (int sft = 16; sft > 0; sft >>= 1) red += __shfl_xor_sync(0xffffffffu, red, sft);
        if (real_tile && lane == 0) pdst[64] = red;
      }
      if (pfe) { long long e2 = clock64(); pe[0] += e1 - e0; pe[1] += e2 - e1; }
    }
    if (pfe) {
      for (int k = 0; k < 2; ++k) prof[27 + k] = pe[k];
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CTA2) cluster_sync_all();   // the peer may still multicast into / arrive on this CTA's barriers
  if (warp == W_MMA) {
    tc_fence_after();
    if constexpr (CTA2)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
  }
}

// CTA pairs are OPT-IN (WMB200_CTA2=1): correct (same tests as the single-CTA kernel) but measured slower on B200,
// 10.8 vs 9.0 ms per launch at 4096 clips: every MMA now needs both SMs' shared memories to be free, the pair
// advances at the pace of its slower half, and the odd CTA's events reach the issuing thread through a relay hop;
// that costs more than the 21 % of operand bytes it saves (profiles/r1_resblock_cta2_phases.json).
static int cta2_clusters(const void *kernel) {
  static int cached = -2;
  if (cached != -2) return cached;
  const char *env = getenv("WMB200_CTA2");
  if (!(env && env[0] == '1')) return cached = 0;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * sm_count(), 1, 1);
  cfg.blockDim = dim3(RB_THREADS, 1, 1);
  cfg.dynamicSmemBytes = RB_SMEM;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kernel, &cfg) != cudaSuccess) { cudaGetLastError(); n = 0; }
  return cached = n;
}

// host_head: HOST rows of the 1x1 head, w[nw][64] then b[nw] (nw = 1: output 0 only; NHEAD == 3: all outputs);
// host_b12: HOST b1[64], b2[64] or null (then the kernel reads the device pointers b1 / b2 into shared memory)
template <int NHEAD>
static int launch_rb(const void *x, const void *w_img, const float *b1, const float *b2, void *y, float *y32, int B,
                     int T, const float *host_head, int nw, const float *host_b12, float *head_out, float *partials,
                     const int *valid_len, cudaStream_t st) {
  using P = typename RbParamsFor<NHEAD>::type;
  static_assert(sizeof(P) <= 8192, "kernel parameter budget");
  P hp;
  memset(&hp, 0, sizeof(hp));
  if (NHEAD > 0) {
    memcpy(hp.w, host_head, sizeof(float) * 64 * nw);
    if (NHEAD == 3) {   // message rows transposed to [channel][bit]: a channel's 16 weights are one 64-byte run
      memset(hp.w + 64, 0, sizeof(float) * 64 * (WM_FUSED_VOTE_MAX - 1));
      for (int j = 1; j < nw; ++j)
        for (int c = 0; c < 64; ++c) hp.w[64 + c * (WM_FUSED_VOTE_MAX - 1) + (j - 1)] = host_head[j * 64 + c];
    }
    memcpy(hp.b, host_head + 64 * nw, sizeof(float) * nw);
  }
  const bool biasp = host_b12 != nullptr;
  if (biasp) {
    memcpy(hp.b1, host_b12, sizeof(float) * 64);
    memcpy(hp.b2, host_b12 + 64, sizeof(float) * 64);
  }
  static bool attr_set = false;
  if (!attr_set) {
    WM_CHECK_CUDA(cudaFuncSetAttribute(resblock_tc_kernel<NHEAD, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, RB_SMEM));
    WM_CHECK_CUDA(cudaFuncSetAttribute(resblock_tc_kernel<NHEAD, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, RB_SMEM));
    WM_CHECK_CUDA(cudaFuncSetAttribute(resblock_tc_kernel<NHEAD, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, RB_SMEM));
    attr_set = true;
  }
  const int nvote = NHEAD == 3 ? nw - 1 : 0;
  long long ntiles = (long long)B * ((T + TO - 1) / TO);
  const int ncl = cta2_clusters(reinterpret_cast<const void *>(resblock_tc_kernel<NHEAD, true, false>));
  if (ncl >= sm_count() / 2 - 2 && ntiles >= 2 && b1 != nullptr && b2 != nullptr) {   // CTA pairs (one per TPC); biases from device memory
    const long long npairs = (ntiles + 1) / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * (unsigned)(npairs < ncl ? npairs : ncl), 1, 1);
    cfg.blockDim = dim3(RB_THREADS, 1, 1);
    cfg.dynamicSmemBytes = RB_SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    WM_CHECK_CUDA(cudaLaunchKernelEx(&cfg, resblock_tc_kernel<NHEAD, true, false>, reinterpret_cast<const uint4 *>(x),
                                     reinterpret_cast<const uint4 *>(w_img), b1, b2, reinterpret_cast<uint4 *>(y), y32, B, T,
                                     hp, nvote, head_out, partials, valid_len, get_profile_buffer()));
    WM_CHECK_LAUNCH("resblock_tc (CTA pairs)");
    return 0;
  }
  int grid = (int)(ntiles < sm_count() ? ntiles : sm_count());
  auto kern = biasp ? resblock_tc_kernel<NHEAD, false, true> : resblock_tc_kernel<NHEAD, false, false>;
  kern<<<grid, RB_THREADS, RB_SMEM, st>>>(reinterpret_cast<const uint4 *>(x), reinterpret_cast<const uint4 *>(w_img), b1, b2,
                                          reinterpret_cast<uint4 *>(y), y32, B, T, hp, nvote, head_out, partials, valid_len,
                                          get_profile_buffer());
  WM_CHECK_LAUNCH("resblock_tc");
  return 0;
}

int launch_resblock_tc(const void *x, const void *w_img, const float *b1, const float *b2, void *y, float *y32, int B,
                       int T, cudaStream_t st, const float *host_b12) {
  if (B == 0 || T == 0) return 0;
  return launch_rb<0>(x, w_img, b1, b2, y, y32, B, T, nullptr, 0, host_b12, nullptr, nullptr, nullptr, st);
}

int resblock_tiles_per_clip(int T) { return (T + TO - 1) / TO; }

// ResBlock + Conv1d(64,1,1): delta_raw[B][T]   (host_head: HOST copy of w[64], b[1])
int launch_resblock_head1_tc(const void *x, const void *w_img, const float *b1, const float *b2,
                             const float *host_head, float *delta_raw, int B, int T, cudaStream_t st,
                             const float *host_b12) {
  if (B == 0 || T == 0) return 0;
  return launch_rb<1>(x, w_img, b1, b2, nullptr, nullptr, B, T, host_head, 1, host_b12, delta_raw, nullptr, nullptr, st);
}

// Detector's last ResBlock + channel 0 of its 1x1 head + sigmoid + per-(tile, warp) partial sums of the
// probability and of the 64 activations; finish with launch_detect_finalize.  host_head: HOST w0[64], b0.
// With host_head_all (HOST w[nout][64] then b[nout], nout <= WM_FUSED_VOTE_MAX) the epilogue also evaluates the
// message logits per sample and counts the positive ones (majority vote of py/main16.py:398).
// partials: B * tiles_per_clip * 4 * WM_DET_PART floats.
int launch_resblock_detect_tc(const void *x, const void *w_img, const float *b1, const float *b2,
                              const float *host_head, const int *valid_len, float *probs, float *partials, int B, int T,
                              cudaStream_t st, const float *host_b12, const float *host_head_all, int nout) {
  if (B == 0 || T == 0) return 0;
  if (host_head_all != nullptr && nout > 1 && nout <= WM_FUSED_VOTE_MAX) {
    float packed[WM_FUSED_VOTE_MAX * 65];   // rows padded to WM_FUSED_VOTE_MAX (missing outputs: zero weights)
    memset(packed, 0, sizeof(packed));
    memcpy(packed, host_head_all, sizeof(float) * 64 * nout);
    memcpy(packed + 64 * WM_FUSED_VOTE_MAX, host_head_all + 64 * nout, sizeof(float) * nout);
    return launch_rb<3>(x, w_img, b1, b2, nullptr, nullptr, B, T, packed, WM_FUSED_VOTE_MAX, host_b12, probs, partials,
                        valid_len, st);
  }
  return launch_rb<2>(x, w_img, b1, b2, nullptr, nullptr, B, T, host_head, 1, host_b12, probs, partials, valid_len, st);
}

// One block per clip: the partial sums are added in a fixed order, then
//   clip_prob[b] = sum(prob) / valid;  msg_logits[b][j] = head_b[1+j] + head_w[1+j] . (sum(activations) / valid)
// (the mean over time of the message logits, py/main16.py:1145-1146; 0 when no sample is valid)
__global__ void __launch_bounds__(128)
    detect_finalize_kernel(const float *__restrict__ partials, const int *__restrict__ valid_len,
                           const float *__restrict__ head_w, const float *__restrict__ head_b,
                           float *__restrict__ clip_prob, float *__restrict__ msg_logits,
                           float *__restrict__ vote_frac, int nparts, int T, int nout) {
  __shared__ float mean_s[WM_DET_PART];
  const int b = blockIdx.x, c = threadIdx.x;
  const int vl = valid_len != nullptr ? min(max(valid_len[b], 0), T) : T;
  if (c <= 64 || (vote_frac != nullptr && c >= WM_DET_VOTE0 && c < WM_DET_VOTE0 + nout - 1)) {
    const float *src = partials + (size_t)b * nparts * WM_DET_PART + c;
    float s = 0.0f;
    for (int i = 0; i < nparts; ++i) s += src[(size_t)i * WM_DET_PART];   // (vote counts: integers, exact in fp32)
    mean_s[c] = vl > 0 ? s / (float)vl : 0.0f;
    if (c >= WM_DET_VOTE0) vote_frac[(size_t)b * (nout - 1) + c - WM_DET_VOTE0] = mean_s[c];
  }
  __syncthreads();
  if (c == 0 && clip_prob) clip_prob[b] = mean_s[64];
  if (c >= 1 && c < nout && msg_logits) {
    float a = vl > 0 ? head_b[c] : 0.0f;
    for (int k = 0; k < 64; ++k) a = fmaf(head_w[c * 64 + k], mean_s[k], a);
    msg_logits[(size_t)b * (nout - 1) + c - 1] = a;
  }
}

int launch_detect_finalize(const float *partials, const int *valid_len, const float *head_w, const float *head_b,
                           float *clip_prob, float *msg_logits, float *vote_frac, int B, int T, int nout, cudaStream_t st) {
  if (B == 0) return 0;
  detect_finalize_kernel<<<B, 128, 0, st>>>(partials, valid_len, head_w, head_b, clip_prob, msg_logits, vote_frac,
                                            4 * ((T + TO - 1) / TO), T, nout);
  WM_CHECK_LAUNCH("detect_finalize");
  return 0;
}

}  // namespace wm
