// Convolutions of the main14b_2 residual stack (py/main14b_2.py:83-224, BASELINE config 3) as implicit GEMMs on the
// 5th-generation tensor cores (sm_100a): strided Conv1d k3, the residual blocks' conv2 with the 1x1 strided skip
// convolution folded in as extra K, ConvTranspose1d(k = 2s, stride s) as a 2- or 3-tap convolution over s * Cout
// phase columns, the k7 output convolution.  Channel counts 16..512.
//
// GEMM view.  Rows (M) = time steps of ALL clips flattened: clip c, step t sits at row c * (T + GAP) + GAP + t of every
// plane; the GAP zero rows in front of a clip are the convolution's zero padding, so a tile of 128 consecutive rows
// may straddle clips.  Columns (N) = output channels (times phases for the transposed convolution) in chunks of
// NC <= 128.  K = input channels x taps, walked in stages of 16 channels: one stage = one 16-channel slice of one
// source tile (128 + taps - 1 rows, the taps being descriptor start addresses 16 bytes apart) and its taps' weights.
//
// Precision: the bf16-pair scheme of wm_conv_tc.cu (v = hi + lo; hi x [W_hi | W_lo] with N = 2 NC and lo x W_hi with
// N = NC into the same accumulator; the epilogue adds column n and NC + n).
//
// Activations ("planar"): per tensor 2 * C/8 planes of plane_rows rows x 16 B; plane g = bf16 hi of channels
// 8g..8g+7, plane C/8 + g = bf16 lo.  A plane is the tcgen05 no-swizzle K-major canonical layout, so operand tiles are
// linear bulk copies.  A producer may write its rows split by phase (row t -> buffer t % s, row t / s): the strided
// convolution that follows then reads contiguous rows (tap 0: phase s-1 one row up, tap 1: phase 0, tap 2: phase 1;
// the 1x1 stride-s skip convolution: phase 0).
#include <cuda_bf16.h>

#include "wm_common.h"
#include "wm_tc.cuh"

namespace wm {

using namespace tc;

namespace {

constexpr int GAP = WM_PC_GAP;
constexpr int TILE = 128;
constexpr int A_PLANE = (TILE + 6) * 16;   // shared-memory pitch of one 8-channel plane of a stage (taps <= 7)
constexpr int A_BYTES = 4 * A_PLANE;       // hi k8 = 0,1 then lo k8 = 0,1
constexpr int MAX_STAGE = 16;

struct KSrc {
  const uint4 *base;
  int kchunks;     // cin / 16
  int lo_plane;    // cin / 8: first lo plane
  int row_off;
  int taps;
};

struct KParams {
  KSrc src[3];
  int nsrc;
  long long plane_rows;        // of the sources and the residual
  int B, T, Tp;                // row geometry
  long long R;                 // B * Tp + GAP rows
  const uint8_t *w;
  long long w_chunk_bytes;     // packed weights of one N chunk
  const float *bias;
  int nch;                     // N chunks
  int n_total;
  int elu;
  const uint4 *residual;
  int mode;
  void *y;
  long long out_plane_rows;
  int out_split;
  long long out_phase_rows;    // rows (16 B units) between phase buffers of the output
  int out_Tp;                  // mode 0 split: T / split + GAP; mode 2: out_T + GAP
  int ct_stride, ct_pad, ct_cout, out_T;
  int ct_pk;                   // phases interleaved per 8-channel group (>= 1)
  int stage_bytes, nstage;
  signed char chunk_off[64];
  // fused residual block (pconv_rb_kernel): second GEMM on the first one's output kept in shared memory
  const uint8_t *w2;           // packed conv2 weights: 16-channel group major, 3 taps each; then the skip source's slices
  const float *bias2;
  KSrc skip;                   // kchunks == 0: none
  int dbg;                     // developer A/B switches (timing experiments only): 1 = no global stores, 2 = no MMAs
  long long *prof;             // developer hook (wm_debug_lstm_profile buffer): per-role cycle sums of block 0, or null
  int kps;                     // fused block: 16-channel slices per stage (2 when every source has an even number)
  int u_off, w1_bytes, w2_bytes;   // shared-memory offsets / sizes set by the launcher (w2_bytes includes the skip slices)
};

// two adjacent 16-byte rows as one 256-bit store (sm_100: STG.256)
__device__ __forceinline__ void st_global_256(uint4 *p, const uint4 &a, const uint4 &b) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x),
               "r"(b.y), "r"(b.z), "r"(b.w)
               : "memory");
}

__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float *v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// ELU for the epilogue: exp(v) - 1 through ex2.approx (absolute error ~1e-7, the rounding of exp(v) itself).  expm1f costs
// ~40 instructions per element and made every layer with a short K loop epilogue-issue bound (128 x NC elements per tile).
__device__ __forceinline__ float elu1(float v) { return v > 0.0f ? v : ex2_approx(v * 1.4426950408889634f) - 1.0f; }
// the same on N values with the multiply and the subtraction packed two per instruction
template <int N>
__device__ __forceinline__ void elu_packed(float *o) {
  const f32x2 l2e = pk2(1.4426950408889634f, 1.4426950408889634f), m1 = pk2(-1.0f, -1.0f);
#pragma unroll
  for (int k = 0; k < N; k += 2) {
    float ta, tb;
    upk2(mul2(pk2(o[k], o[k + 1]), l2e), ta, tb);
    float ea, eb;
    upk2(add2(pk2(ex2_approx(ta), ex2_approx(tb)), m1), ea, eb);
    o[k] = o[k] > 0.0f ? o[k] : ea;
    o[k + 1] = o[k + 1] > 0.0f ? o[k + 1] : eb;
  }
}

// Planar store of CS channels (first GEMM column n0) of row m = (clip c, gap-relative row r, step t): same geometry, or
// split by phase for a strided consumer.  Gap rows are written as zeros (they are the next convolution's padding).
template <int CS>
__device__ __forceinline__ void store_planar(const KParams &P, const float *o, int n0, long long m, int c, int r, int t,
                                             bool real) {
  uint4 *y = reinterpret_cast<uint4 *>(P.y);
  const int lo0 = P.n_total >> 3;
  const uint4 z = make_uint4(0, 0, 0, 0);
  if (P.out_split <= 1) {
#pragma unroll
    for (int g = 0; g < CS / 8; ++g) {
      const int pl = (n0 >> 3) + g;
      uint4 hi = z, lo = z;
      if (real) split8(o + g * 8, hi, lo);
      y[(long long)pl * P.out_plane_rows + m] = hi;
      y[(long long)(lo0 + pl) * P.out_plane_rows + m] = lo;
    }
  } else {
    const int sp = P.out_split;
    if (real) {
      const int qq = t / sp, phs = t - qq * sp;
      if (phs <= 1 || phs == sp - 1) {
        uint4 *yp = y + (long long)phs * P.out_phase_rows;
        const long long row = (long long)c * P.out_Tp + GAP + qq;
#pragma unroll
        for (int g = 0; g < CS / 8; ++g) {
          const int pl = (n0 >> 3) + g;
          uint4 hi, lo;
          split8(o + g * 8, hi, lo);
          yp[(long long)pl * P.out_plane_rows + row] = hi;
          yp[(long long)(lo0 + pl) * P.out_plane_rows + row] = lo;
        }
      }
    } else {
      // gap row r of clip c (or of the closing gap): zero the same gap row of every phase buffer that is read
      const long long row = (long long)c * P.out_Tp + r;
      for (int phs = 0; phs < sp; ++phs) {
        if (!(phs <= 1 || phs == sp - 1)) continue;
        uint4 *yp = y + (long long)phs * P.out_phase_rows;
#pragma unroll
        for (int g = 0; g < CS / 8; ++g) {
          const int pl = (n0 >> 3) + g;
          yp[(long long)pl * P.out_plane_rows + row] = z;
          yp[(long long)(lo0 + pl) * P.out_plane_rows + row] = z;
        }
      }
    }
  }
}

template <int NC>
struct Cfg {
  static constexpr int NSL = NC >= 32 ? 4 : 2;    // column slices of the epilogue
  static constexpr int CS = NC / NSL;             // channels per slice: 32, 16, 8, 8
  static constexpr int B_TAP = 2 * NC * 32;       // bytes of one tap's [W_hi | W_lo] x 16 channels
  static constexpr int TMEM_COLS = 4 * NC < 32 ? 32 : 4 * NC;   // two accumulators of 2 NC columns
};

// MMAs of one stage: TAPS row-shifted A descriptors against TAPS consecutive weight tiles, all offsets compile-time.
// (A rolled tap loop with run-time descriptor arithmetic cost ~30 uniform-datapath instructions per tap on the single
// issuing thread — more than the 88..112 tensor cycles a tap of a 32/64-column chunk takes: the layers with few
// channels were bound by MMA ISSUE, tensor pipe 42 % busy and never waiting on a barrier.)
template <int TAPS, int NC>
__device__ __forceinline__ void issue_taps(uint32_t d_tmem, uint64_t a0, uint64_t b0, uint32_t accum) {
  constexpr uint32_t idesc_hi = make_idesc(128, 2 * NC), idesc_lo = make_idesc(128, NC);
#pragma unroll
  for (int tp = 0; tp < TAPS; ++tp) {
    mma_bf16(d_tmem, a0 + (uint64_t)tp, b0 + (uint64_t)(tp * ((2 * NC * 32) >> 4)), idesc_hi, tp == 0 ? accum : 1u);
    mma_bf16(d_tmem, a0 + (uint64_t)(tp + ((2 * A_PLANE) >> 4)), b0 + (uint64_t)(tp * ((2 * NC * 32) >> 4)), idesc_lo, 1u);
  }
}
template <int NC>
__device__ __forceinline__ void issue_stage(int taps, uint32_t d_tmem, uint64_t a0, uint64_t b0, uint32_t accum) {
  switch (taps) {
    case 1: issue_taps<1, NC>(d_tmem, a0, b0, accum); break;
    case 2: issue_taps<2, NC>(d_tmem, a0, b0, accum); break;
    case 3: issue_taps<3, NC>(d_tmem, a0, b0, accum); break;
    case 7: issue_taps<7, NC>(d_tmem, a0, b0, accum); break;
    default:
      for (int tp = 0; tp < taps; ++tp) issue_taps<1, NC>(d_tmem, a0 + (uint64_t)tp, b0 + (uint64_t)(tp * ((2 * NC * 32) >> 4)), tp == 0 ? accum : 1u);
  }
}

// Warps 0..15 epilogue (TMEM lane quadrant = warp % 4, column slice = warp / 4), warp 16 producer, warp 17 MMA issuer.
template <int NC>
__global__ void __launch_bounds__(576, 1) pconv_tc_kernel(const __grid_constant__ KParams P) {
  using C = Cfg<NC>;
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t s_base = smem_u32(smem);
  const int NS = P.nstage;
  const uint32_t bars = s_base + NS * P.stage_bytes;
  auto full_bar = [&](int s) { return bars + 8 * s; };
  auto empty_bar = [&](int s) { return bars + 8 * (MAX_STAGE + s); };
  auto tfull_bar = [&](int a) { return bars + 8 * (2 * MAX_STAGE + a); };
  auto tempty_bar = [&](int a) { return bars + 8 * (2 * MAX_STAGE + 2 + a); };
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + NS * P.stage_bytes + 8 * (2 * MAX_STAGE + 4));
  // the layer's biases, staged once (per-tile __ldg of 8 float4 sat on the epilogue warps' critical path)
  float *bias_s = reinterpret_cast<float *>(smem + NS * P.stage_bytes + 8 * (2 * MAX_STAGE + 4) + 16);
  for (int i = threadIdx.x; i < P.n_total; i += blockDim.x) bias_s[i] = P.bias[i];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long mtiles = (P.R + TILE - 1) / TILE;
  const long long ntiles = mtiles * P.nch;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 4 * C::NSL); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 17) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(C::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 16) {
    // ===== producer: per stage 4 plane copies (hi, hi, lo, lo of one 16-channel slice) + the taps' weights =====
    // ONE ELECTED lane walks the pipeline and issues the five copies of a stage; the other lanes of the warp idle.
    // History of this role (cycles of producer time per stage, whatever the stage held): one thread under `lane == 0`
    // with 64-bit index arithmetic ~1000; one copy per lane under `lane < 5` ~550 (every cp.async.bulk sat in its own
    // operand-uniformising loop); an elected lane ~400 (elect.sync tells ptxas exactly one lane is active: the copies
    // take their operands from uniform registers) and ~270 once the other lanes stopped polling along.  One TMA tensor
    // load per stage was tried instead and is slower: a box with a 16-byte inner dimension moves 16 bytes per request.
    uint32_t s = 0, ph = 0;
    const uint32_t nch = (uint32_t)P.nch;
    const bool leader = elect_one();
    const long long prow = P.plane_rows;
    for (uint32_t tile = blockIdx.x; leader && tile < (uint32_t)ntiles; tile += gridDim.x) {
      const uint32_t mt = tile / nch, j = tile - mt * nch;
      const long long m0 = (long long)mt * TILE;
      const uint8_t *wj = P.w + (size_t)j * P.w_chunk_bytes;
      for (int si = 0; si < P.nsrc; ++si) {
        const KSrc &S = P.src[si];
        const uint32_t abytes = (uint32_t)(TILE + S.taps - 1) * 16u;
        const uint32_t wbytes = (uint32_t)S.taps * C::B_TAP;
        const uint4 *hi = S.base + (m0 + S.row_off + (si == 0 ? (int)P.chunk_off[j] : 0));
        const uint4 *lo = hi + (long long)S.lo_plane * prow;
        const int kch = S.kchunks;
        for (int kc = 0; kc < kch; ++kc) {
          mbar_wait(empty_bar(s), ph ^ 1);
          {
            const uint32_t dst = s_base + s * P.stage_bytes, fb = full_bar(s);
            mbar_arrive_expect_tx(fb, 4 * abytes + wbytes);
            bulk_g2s(dst, hi, abytes, fb);
            bulk_g2s(dst + A_PLANE, hi + prow, abytes, fb);
            bulk_g2s(dst + 2 * A_PLANE, lo, abytes, fb);
            bulk_g2s(dst + 3 * A_PLANE, lo + prow, abytes, fb);
            bulk_g2s(dst + A_BYTES, wj, wbytes, fb);
          }
          hi += 2 * prow;
          lo += 2 * prow;
          wj += wbytes;
          if (++s == (uint32_t)NS) { s = 0; ph ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 17) {
    // ===== MMA issuer =====
    const bool issuer = elect_one();
    uint32_t s = 0, ph = 0, a = 0, aph = 0;
    // running descriptors of the current stage: A tile, and its weights (a compile-time distance behind it)
    const uint64_t a_base = smem_desc(s_base, A_PLANE, 128);
    const uint64_t b_delta = smem_desc(A_BYTES, 2 * NC * 16, 128) - smem_desc(0, A_PLANE, 128);
    const uint64_t st_step = (uint64_t)(P.stage_bytes >> 4);
    uint64_t a_cur = a_base;
    for (uint32_t tile = blockIdx.x; tile < (uint32_t)ntiles; tile += gridDim.x) {
      mbar_wait_warp(tempty_bar(a), aph ^ 1);
      const uint32_t d_tmem = tmem_base + a * (2 * NC);
      uint32_t accum = 0;
      for (int si = 0; si < P.nsrc; ++si) {
        const int taps = P.src[si].taps, kch = P.src[si].kchunks;
        for (int kc = 0; kc < kch; ++kc) {
          mbar_wait_warp(full_bar(s), ph);
          tc_fence_after();
          if (issuer) {
            if (!(P.dbg & 2)) issue_stage<NC>(taps, d_tmem, a_cur, a_cur + b_delta, accum);
            tc_commit(empty_bar(s));
          }
          accum = 1u;
          __syncwarp();
          a_cur += st_step;
          if (++s == (uint32_t)NS) { s = 0; ph ^= 1; a_cur = a_base; }
        }
      }
      if (issuer) tc_commit(tfull_bar(a));
      __syncwarp();
      a ^= 1;
      if (a == 0) aph ^= 1;
    }
  } else if ((warp >> 2) < C::NSL) {
    // ===== epilogue =====
    const int q = warp & 3, p = warp >> 2;
    constexpr int CS = C::CS;
    const int Tp = P.Tp;
    uint32_t i = 0;
    const uint32_t nch = (uint32_t)P.nch;
    for (uint32_t tile = blockIdx.x; tile < (uint32_t)ntiles; tile += gridDim.x, ++i) {
      const uint32_t a = i & 1u;
      const uint32_t aph = (i >> 1) & 1u;
      const uint32_t mt = tile / nch;
      const int j = (int)(tile - mt * nch);
      const long long m = (long long)mt * TILE + q * 32 + lane;
      const int c = (int)((uint32_t)m / (uint32_t)Tp);
      const int r = (int)(m - (long long)c * Tp);
      const int t = r - GAP;
      const bool inrange = m < P.R;
      const int n0 = j * NC + p * CS;        // first GEMM column of this thread's slice
      if (P.residual != nullptr && inrange && c < P.B && t >= 0) {
        const int lo0 = P.n_total >> 3;      // residual rows towards L1 while the accumulator is still being computed
#pragma unroll
        for (int g = 0; g < CS / 8; ++g) {
          const int pl = (n0 >> 3) + g;
          prefetch_l1(P.residual + ((long long)pl * P.plane_rows + m));
          prefetch_l1(P.residual + ((long long)(lo0 + pl) * P.plane_rows + m));
        }
      }

      mbar_wait_warp(tfull_bar(a), aph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + a * (2 * NC) + ((uint32_t)(q * 32) << 16) + p * CS;
      float o[CS];
      {
        float v2[CS];
#pragma unroll
        for (int g = 0; g < CS / 8; ++g) {
          tmem_ld8(taddr + g * 8, o + g * 8);
          tmem_ld8(taddr + NC + g * 8, v2 + g * 8);
        }
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive_warp(tempty_bar(a));
        add_packed<CS>(o, v2);
      }
      if (!inrange) continue;
      {
        const float4 *bp = reinterpret_cast<const float4 *>(bias_s + n0);
#pragma unroll
        for (int k = 0; k < CS / 4; ++k) {
          const float4 b = bp[k];
          add_packed<4>(o + 4 * k, &b.x);
        }
      }
      const bool real = (c < P.B) && (t >= 0);
      if (P.residual != nullptr && real) {
        const int lo0 = P.n_total >> 3;
#pragma unroll
        for (int g = 0; g < CS / 8; ++g) {
          const int pl = (n0 >> 3) + g;
          const uint4 rh = __ldg(P.residual + ((long long)pl * P.plane_rows + m));
          const uint4 rl = __ldg(P.residual + ((long long)(lo0 + pl) * P.plane_rows + m));
          float rr[8];
          join8(rh, rl, rr);
          add_packed<8>(o + g * 8, rr);
        }
      }
      if (P.elu) elu_packed<CS>(o);
      if (P.dbg & 1) {                       // timing experiment: everything but the stores
        float acc = 0.0f;
#pragma unroll
        for (int k = 0; k < CS; ++k) acc += o[k];
        if (acc == 1.2345e-30f) reinterpret_cast<float *>(P.y)[0] = acc;
        continue;
      }

      if (P.mode == WM_PC_OUT_PLANAR) {
        store_planar<CS>(P, o, n0, m, c, r, t, real);
      } else if (P.mode == WM_PC_OUT_CONVT) {
        // column n = (phase, co) (see wm_pconv.ct_interleave); row (c, q = t) -> output step s q + phase of clip c.  The
        // first gap row of a clip also carries the last outputs of the previous clip when out_T > s * T (odd strides).
        uint4 *y = reinterpret_cast<uint4 *>(P.y);
        const int s = P.ct_stride, cout = P.ct_cout, lo0 = cout >> 3, pk = P.ct_pk;
        const uint4 z = make_uint4(0, 0, 0, 0);
        auto decode = [&](int n, int &phs, int &pl) {      // 8-column group starting at n -> phase, 8-channel plane
          const int blk = cout * pk, kb = n / blk, rem = n - kb * blk, gg = rem / (8 * pk);
          phs = kb * pk + ((rem - gg * 8 * pk) >> 3);
          pl = gg;
        };
        auto store_one = [&](int g) {
          int phs, pl;
          decode(n0 + g * 8, phs, pl);
          const int tout = s * t + phs;
          if (tout >= 0) {
            if (c < P.B && tout < P.out_T) {
              uint4 hi, lo;
              split8(o + g * 8, hi, lo);
              const long long row = (long long)c * P.out_Tp + GAP + tout;
              y[(long long)pl * P.out_plane_rows + row] = hi;
              y[(long long)(lo0 + pl) * P.out_plane_rows + row] = lo;
            }
          } else {
            if (tout >= -GAP) {
              const long long row = (long long)c * P.out_Tp + GAP + tout;
              y[(long long)pl * P.out_plane_rows + row] = z;
              y[(long long)(lo0 + pl) * P.out_plane_rows + row] = z;
            }
            if (r == 0 && c >= 1) {
              const int tout2 = s * P.T + phs;
              if (tout2 < P.out_T) {
                uint4 hi, lo;
                split8(o + g * 8, hi, lo);
                const long long row = (long long)(c - 1) * P.out_Tp + GAP + tout2;
                y[(long long)pl * P.out_plane_rows + row] = hi;
                y[(long long)(lo0 + pl) * P.out_plane_rows + row] = lo;
              }
            }
          }
        };
        if constexpr (CS >= 16) {
          // with an even interleave two consecutive 8-column groups of a thread are phases ph, ph + 1 of the same 8
          // channels = two adjacent 16-byte rows: ONE 32-byte store each for hi and lo.  (Rows a stride apart made every
          // 16-byte store its own half-empty sector: 4096 sector writes per tile, the stride-4 layer's bottleneck.)
#pragma unroll
          for (int g = 0; g < CS / 8; g += 2) {
            int phs, pl;
            decode(n0 + g * 8, phs, pl);
            const int tout = s * t + phs;
            const long long row = (long long)c * P.out_Tp + GAP + tout;
            if ((pk & 1) == 0 && (phs & 1) == 0 && c < P.B && tout >= 0 && tout + 1 < P.out_T && (row & 1) == 0) {
              uint4 h0, l0, h1, l1;
              split8(o + g * 8, h0, l0);
              split8(o + g * 8 + 8, h1, l1);
              st_global_256(y + ((long long)pl * P.out_plane_rows + row), h0, h1);
              st_global_256(y + ((long long)(lo0 + pl) * P.out_plane_rows + row), l0, l1);
            } else {
              store_one(g);
              store_one(g + 1);
            }
          }
        } else {
#pragma unroll
          for (int g = 0; g < CS / 8; ++g) store_one(g);
        }
      } else {
        // fp32 channels-first y[c][ch][t], ch < ct_cout, t < out_T
        if (real && t < P.out_T) {
          float *y = reinterpret_cast<float *>(P.y);
#pragma unroll
          for (int k = 0; k < CS; ++k) {
            const int ch = n0 + k;
            if (ch < P.ct_cout) y[((long long)c * P.ct_cout + ch) * P.out_T + t] = o[k];
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 17) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C::TMEM_COLS) : "memory");
  }
}

template <int NC>
int launch_pconv_t(const KParams &P0, cudaStream_t st) {
  using C = Cfg<NC>;
  KParams P = P0;
  int maxtaps = 1;
  for (int i = 0; i < P.nsrc; ++i) maxtaps = P.src[i].taps > maxtaps ? P.src[i].taps : maxtaps;
  P.stage_bytes = (A_BYTES + maxtaps * C::B_TAP + 127) / 128 * 128;
  const int budget = 200 * 1024;   // + barriers and up to 8 KB of biases
  int ns = budget / P.stage_bytes;
  ns = ns > MAX_STAGE ? MAX_STAGE : ns;
  WM_CHECK_ARG(ns >= 2, "pconv: a stage of %d bytes does not fit the shared memory twice", P.stage_bytes);
  P.nstage = ns;
  const int smem_bytes = ns * P.stage_bytes + 8 * (2 * MAX_STAGE + 4) + 16 + P.n_total * 4;
  static int attr_bytes = 0;
  if (smem_bytes > attr_bytes) {
    WM_CHECK_CUDA(cudaFuncSetAttribute(pconv_tc_kernel<NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_bytes = 227 * 1024;
  }
  const long long ntiles = ((P.R + TILE - 1) / TILE) * P.nch;
  WM_CHECK_ARG(P.R < (1LL << 31) && ntiles < (1LL << 31), "pconv: %lld rows x %d chunks exceed the 32-bit tile index", P.R,
               P.nch);
  const int grid = (int)(ntiles < sm_count() ? ntiles : sm_count());
  pconv_tc_kernel<NC><<<grid, 576, smem_bytes, st>>>(P);
  WM_CHECK_LAUNCH("pconv_tc");
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// Fused ResidualBlock (py/main14b_2.py:97-105) for blocks of up to 64 output channels, the ones that run at
// T = 2000..16000 and are bound by HBM when conv1's output makes a round trip:
//     u = elu(conv1(x) + b1)  ->  bf16-pair tile in SHARED MEMORY  ->  y = elu(conv2(u) + b2 + skip(x) | x)
// GEMM 1 is the stage pipeline of pconv_tc_kernel (any source list: plain k3, or the three phase sources of a strided
// conv1); its epilogue writes the 128-row u tile in the canonical operand layout; GEMM 2 reads it with three row-shifted
// descriptors against conv2's weights RESIDENT in shared memory, optionally followed by streamed stages of the 1x1
// skip convolution (phase 0 of x).  A tile yields 126 output rows (u rows u0 .. u0 + 127 -> y rows u0 + 1 .. u0 + 126).
// The identity residual is read from global memory in the second epilogue (an L2 hit: the same rows were just loaded).
// Warps 0..7 epilogue 1, 8..15 epilogue 2 (quadrant = warp % 4, half of the channels = (warp / 4) % 2), 16 producer,
// 17 MMA issuer.  The issuer runs GEMM 1 of tile i + 1 before GEMM 2 of tile i, so epilogue 1 overlaps tensor work.
// ---------------------------------------------------------------------------------------------------------------
#define RB_TICK(k) do { if (pf) { const long long now_ = clock64(); pa[k] += now_ - last_; last_ = now_; } } while (0)
constexpr int RB_ROWS = 126;
constexpr int U_PLANE = 136 * 16;

template <int NC>
__global__ void __launch_bounds__(576, 1) pconv_rb_kernel(const __grid_constant__ KParams P) {
  using C = Cfg<NC>;
  constexpr int CS = NC / 2;                 // channels per epilogue thread
  constexpr int U_BYTES = 2 * (NC / 8) * U_PLANE;
  constexpr int KC2 = NC / 16;               // 16-channel groups of u
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t s_base = smem_u32(smem);
  const int NS = P.nstage;
  const uint32_t w1_smem = s_base + NS * P.stage_bytes;     // BOTH convolutions' weights stay in shared memory: streaming
  const uint32_t w2_smem = w1_smem + P.w1_bytes;            // conv1's with every stage cost more L2 traffic than the
  const uint32_t u_smem = s_base + P.u_off;                 // activations and left room for only 5 stages of prefetch
  const uint32_t bars = u_smem + 2 * U_BYTES;
  auto full_bar = [&](int s) { return bars + 8 * s; };
  auto empty_bar = [&](int s) { return bars + 8 * (MAX_STAGE + s); };
  const uint32_t xb = bars + 8 * 2 * MAX_STAGE;
  auto t1_full = [&](int a) { return xb + 8 * a; };
  auto t1_empty = [&](int a) { return xb + 8 * (2 + a); };
  auto u_full = [&](int a) { return xb + 8 * (4 + a); };
  auto u_empty = [&](int a) { return xb + 8 * (6 + a); };
  auto t2_full = [&](int a) { return xb + 8 * (8 + a); };
  auto t2_empty = [&](int a) { return xb + 8 * (10 + a); };
  const uint32_t w2bar = xb + 8 * 12;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + P.u_off + 2 * U_BYTES + 8 * (2 * MAX_STAGE + 13));
  float *bias_s = reinterpret_cast<float *>(smem + P.u_off + 2 * U_BYTES + 8 * (2 * MAX_STAGE + 13) + 24);   // b1 then b2 (16-byte aligned)
  if (threadIdx.x < 2 * NC) bias_s[threadIdx.x] = threadIdx.x < NC ? P.bias[threadIdx.x] : P.bias2[threadIdx.x - NC];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t ntiles = (uint32_t)((P.R + RB_ROWS - 1) / RB_ROWS);
  const uint32_t nmine = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) {
      mbar_init(t1_full(a), 1); mbar_init(t1_empty(a), 8);
      mbar_init(u_full(a), 8); mbar_init(u_empty(a), 1);
      mbar_init(t2_full(a), 1); mbar_init(t2_empty(a), 8);
    }
    mbar_init(w2bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 17) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(8 * NC)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t acc1 = tmem_base, acc2 = tmem_base + 4 * NC;

  if (warp == 16) {
    // ===== producer =====
    if (lane == 0) {
      mbar_arrive_expect_tx(w2bar, (uint32_t)(P.w1_bytes + P.w2_bytes));
      for (int off = 0; off < P.w1_bytes; off += C::B_TAP) bulk_g2s(w1_smem + off, P.w + off, C::B_TAP, w2bar);
      for (int off = 0; off < P.w2_bytes; off += C::B_TAP) bulk_g2s(w2_smem + off, P.w2 + off, C::B_TAP, w2bar);
    }
    uint32_t s = 0, ph = 0;
    const bool leader = elect_one();
    const bool pf = P.prof != nullptr && blockIdx.x == 0 && leader;
    long long pa[2] = {0, 0}, last_ = pf ? clock64() : 0;
    const long long prow = P.plane_rows;
    for (uint32_t i = 0; leader && i <= nmine; ++i) {      // only the elected lane walks the pipeline
      if (i < nmine) {
        const long long u0 = (long long)(blockIdx.x + i * gridDim.x) * RB_ROWS - 1;
        for (int si = 0; si < P.nsrc; ++si) {
          const KSrc &S = P.src[si];
          const uint32_t abytes = (uint32_t)(TILE + S.taps - 1) * 16u;
          const uint4 *hi = S.base + (u0 + S.row_off);
          const uint4 *lo = hi + (long long)S.lo_plane * prow;
          const int kch = S.kchunks;
          for (int kc = 0; kc < kch; kc += P.kps) {
            RB_TICK(1);
            mbar_wait(empty_bar(s), ph ^ 1);
            RB_TICK(0);
            const uint32_t fb = full_bar(s);
            mbar_arrive_expect_tx(fb, (uint32_t)P.kps * 4 * abytes);
            for (int q2 = 0; q2 < P.kps; ++q2) {
              const uint32_t dst = s_base + s * P.stage_bytes + q2 * A_BYTES;
              bulk_g2s(dst, hi, abytes, fb);
              bulk_g2s(dst + A_PLANE, hi + prow, abytes, fb);
              bulk_g2s(dst + 2 * A_PLANE, lo, abytes, fb);
              bulk_g2s(dst + 3 * A_PLANE, lo + prow, abytes, fb);
              hi += 2 * prow;
              lo += 2 * prow;
            }
            if (++s == (uint32_t)NS) { s = 0; ph ^= 1; }
          }
        }
      }
      if (i >= 1 && P.skip.kchunks > 0) {
        // skip stages of tile i - 1: phase-0 rows of x under the tile's OUTPUT rows (u0 + 1 ..)
        const KSrc &S = P.skip;
        const uint32_t abytes = (uint32_t)TILE * 16u;
        const uint4 *hi = S.base + ((long long)(blockIdx.x + (i - 1) * gridDim.x) * RB_ROWS + S.row_off);
        const uint4 *lo = hi + (long long)S.lo_plane * prow;
        for (int kc = 0; kc < S.kchunks; kc += P.kps) {
          mbar_wait(empty_bar(s), ph ^ 1);
          const uint32_t fb = full_bar(s);
          mbar_arrive_expect_tx(fb, (uint32_t)P.kps * 4 * abytes);
          for (int q2 = 0; q2 < P.kps; ++q2) {
            const uint32_t dst = s_base + s * P.stage_bytes + q2 * A_BYTES;
            bulk_g2s(dst, hi, abytes, fb);
            bulk_g2s(dst + A_PLANE, hi + prow, abytes, fb);
            bulk_g2s(dst + 2 * A_PLANE, lo, abytes, fb);
            bulk_g2s(dst + 3 * A_PLANE, lo + prow, abytes, fb);
            hi += 2 * prow;
            lo += 2 * prow;
          }
          if (++s == (uint32_t)NS) { s = 0; ph ^= 1; }
        }
      }
    }
    if (pf) { P.prof[0] = pa[0]; P.prof[1] = pa[1]; P.prof[31] = nmine; }
    __syncwarp();
  } else if (warp == 17) {
    // ===== MMA issuer =====
    const bool issuer = elect_one();
    constexpr uint32_t idesc_hi = make_idesc(128, 2 * NC), idesc_lo = make_idesc(128, NC);
    uint32_t s = 0, ph = 0;
    const bool pf = P.prof != nullptr && blockIdx.x == 0 && issuer;
    long long pa[8] = {0, 0, 0, 0, 0, 0, 0, 0}, last_ = pf ? clock64() : 0;
    const uint64_t a_base = smem_desc(s_base, A_PLANE, 128), st_step = (uint64_t)(P.stage_bytes >> 4);
    const uint64_t w1_desc = smem_desc(w1_smem, 2 * NC * 16, 128), w2_desc = smem_desc(w2_smem, 2 * NC * 16, 128);
    const uint64_t u_desc = smem_desc(u_smem, U_PLANE, 128);
    uint64_t a_cur = a_base;
    mbar_wait_warp(w2bar, 0);
    for (uint32_t i = 0; i <= nmine; ++i) {
      if (i < nmine) {
        const uint32_t a = i & 1u, aph = (i >> 1) & 1u;
        RB_TICK(7);
        mbar_wait_warp(t1_empty(a), aph ^ 1);
        RB_TICK(0);
        const uint32_t d_tmem = acc1 + a * (2 * NC);
        uint32_t accum = 0;
        uint64_t bd = w1_desc;                                   // walks conv1's resident slices in stage order
        for (int si = 0; si < P.nsrc; ++si) {
          const int taps = P.src[si].taps, kch = P.src[si].kchunks;
          for (int kc = 0; kc < kch; kc += P.kps) {
            RB_TICK(7);
            mbar_wait_warp(full_bar(s), ph);
            RB_TICK(1);
            tc_fence_after();
            if (issuer) {
              issue_stage<NC>(taps, d_tmem, a_cur, bd, accum);
              if (P.kps == 2)
                issue_stage<NC>(taps, d_tmem, a_cur + (uint64_t)(A_BYTES >> 4), bd + (uint64_t)(taps * (C::B_TAP >> 4)), 1u);
              tc_commit(empty_bar(s));
            }
            RB_TICK(2);
            accum = 1u;
            bd += (uint64_t)(P.kps * taps * (C::B_TAP >> 4));
            __syncwarp();
            a_cur += st_step;
            if (++s == (uint32_t)NS) { s = 0; ph ^= 1; a_cur = a_base; }
          }
        }
        if (issuer) tc_commit(t1_full(a));
        __syncwarp();
      }
      if (i >= 1) {
        const uint32_t j = i - 1, a = j & 1u, aph = (j >> 1) & 1u;
        RB_TICK(7);
        mbar_wait_warp(u_full(a), aph);
        RB_TICK(3);
        mbar_wait_warp(t2_empty(a), aph ^ 1);
        RB_TICK(4);
        tc_fence_after();
        const uint32_t d_tmem = acc2 + a * (2 * NC);
        if (issuer) {
          const uint64_t a0 = u_desc + (uint64_t)(a * (U_BYTES >> 4));
#pragma unroll
          for (int kc = 0; kc < KC2; ++kc) {
#pragma unroll
            for (int tp = 0; tp < 3; ++tp) {
              const uint64_t ad = a0 + (uint64_t)(((2 * kc * U_PLANE) >> 4) + tp);
              const uint64_t bd = w2_desc + (uint64_t)(((kc * 3 + tp) * C::B_TAP) >> 4);
              mma_bf16(d_tmem, ad, bd, idesc_hi, (kc | tp) == 0 ? 0u : 1u);
              mma_bf16(d_tmem, ad + (uint64_t)(((NC / 8) * U_PLANE) >> 4), bd, idesc_lo, 1u);
            }
          }
        }
        __syncwarp();
        for (int kc = 0; kc < P.skip.kchunks; kc += P.kps) {
          mbar_wait_warp(full_bar(s), ph);
          tc_fence_after();
          if (issuer) {
            issue_taps<1, NC>(d_tmem, a_cur, w2_desc + (uint64_t)(((KC2 * 3 + kc) * C::B_TAP) >> 4), 1u);
            if (P.kps == 2)
              issue_taps<1, NC>(d_tmem, a_cur + (uint64_t)(A_BYTES >> 4), w2_desc + (uint64_t)(((KC2 * 3 + kc + 1) * C::B_TAP) >> 4), 1u);
            tc_commit(empty_bar(s));
          }
          __syncwarp();
          a_cur += st_step;
          if (++s == (uint32_t)NS) { s = 0; ph ^= 1; a_cur = a_base; }
        }
        if (issuer) {
          tc_commit(u_empty(a));
          tc_commit(t2_full(a));
        }
        RB_TICK(5);
        __syncwarp();
      }
    }
    if (pf) { for (int k = 0; k < 8; ++k) P.prof[8 + k] = pa[k]; }
  } else if (warp < 8) {
    // ===== epilogue 1: accumulator -> elu -> bf16 pair -> u tile in shared memory =====
    const int q = warp & 3, p = warp >> 2;
    const int n0 = p * CS;
    const int Tp = P.Tp;
    const bool pf = P.prof != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
    long long pa[6] = {0, 0, 0, 0, 0, 0}, last_ = pf ? clock64() : 0;
    for (uint32_t i = 0; i < nmine; ++i) {
      const uint32_t a = i & 1u, aph = (i >> 1) & 1u;
      const long long mu = (long long)(blockIdx.x + i * gridDim.x) * RB_ROWS - 1 + q * 32 + lane;
      bool real = false;
      if (mu >= 0 && mu < P.R) {
        const int c = (int)((uint32_t)mu / (uint32_t)Tp);
        real = c < P.B && (int)(mu - (long long)c * Tp) >= GAP;
      }
      RB_TICK(5);
      mbar_wait_warp(t1_full(a), aph);
      RB_TICK(0);
      tc_fence_after();
      const uint32_t taddr = acc1 + a * (2 * NC) + ((uint32_t)(q * 32) << 16) + n0;
      float o[CS];
      {
        float v2[CS];
#pragma unroll
        for (int g = 0; g < CS / 8; ++g) {
          tmem_ld8(taddr + g * 8, o + g * 8);
          tmem_ld8(taddr + NC + g * 8, v2 + g * 8);
        }
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive_warp(t1_empty(a));
        add_packed<CS>(o, v2);
      }
      {
        const float4 *bp = reinterpret_cast<const float4 *>(bias_s + n0);
#pragma unroll
        for (int k = 0; k < CS / 4; ++k) {
          const float4 b = bp[k];
          add_packed<4>(o + 4 * k, &b.x);
        }
      }
      RB_TICK(1);
      mbar_wait_warp(u_empty(a), aph ^ 1);        // GEMM 2 of tile i - 2 has read this buffer
      RB_TICK(2);
      uint8_t *ub = smem + P.u_off + a * U_BYTES + (q * 32 + lane) * 16;
#pragma unroll
      for (int g = 0; g < CS / 8; ++g) {
        uint4 hi = make_uint4(0, 0, 0, 0), lo = hi;
        if (real) {
          elu_packed<8>(o + g * 8);
          split8(o + g * 8, hi, lo);
        }
        const int pl = (n0 >> 3) + g;
        *reinterpret_cast<uint4 *>(ub + pl * U_PLANE) = hi;
        *reinterpret_cast<uint4 *>(ub + (NC / 8 + pl) * U_PLANE) = lo;
      }
      RB_TICK(3);
      fence_async_smem();                          // generic-proxy stores -> the MMA's operand reads
      mbar_arrive_warp(u_full(a));
      RB_TICK(4);
    }
    if (pf) { for (int k = 0; k < 6; ++k) P.prof[16 + k] = pa[k]; }
  } else if (warp < 16) {
    // ===== epilogue 2: accumulator + bias (+ residual) -> elu -> planar output =====
    const int q = warp & 3, p = (warp >> 2) & 1;
    const int n0 = p * CS;
    const int Tp = P.Tp;
    const bool pf = P.prof != nullptr && blockIdx.x == 0 && threadIdx.x == 256;
    long long pa[4] = {0, 0, 0, 0}, last_ = pf ? clock64() : 0;
    for (uint32_t i = 0; i < nmine; ++i) {
      const uint32_t a = i & 1u, aph = (i >> 1) & 1u;
      const int ri = q * 32 + lane;
      const long long m = (long long)(blockIdx.x + i * gridDim.x) * RB_ROWS + ri;
      const int c = (int)((uint32_t)m / (uint32_t)Tp);
      const int r = (int)(m - (long long)c * Tp);
      const int t = r - GAP;
      const bool real = (c < P.B) && (t >= 0) && ri < RB_ROWS;
      if (P.residual != nullptr && real) {
        // start the residual rows towards L1 now: their latency then overlaps the wait for the accumulator
        const int lo0 = P.n_total >> 3;
#pragma unroll
        for (int g = 0; g < CS / 8; ++g) {
          const int pl = (n0 >> 3) + g;
          prefetch_l1(P.residual + ((long long)pl * P.plane_rows + m));
          prefetch_l1(P.residual + ((long long)(lo0 + pl) * P.plane_rows + m));
        }
      }
      RB_TICK(3);
      mbar_wait_warp(t2_full(a), aph);
      RB_TICK(0);
      tc_fence_after();
      const uint32_t taddr = acc2 + a * (2 * NC) + ((uint32_t)(q * 32) << 16) + n0;
      float o[CS];
      {
        float v2[CS];
#pragma unroll
        for (int g = 0; g < CS / 8; ++g) {
          tmem_ld8(taddr + g * 8, o + g * 8);
          tmem_ld8(taddr + NC + g * 8, v2 + g * 8);
        }
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive_warp(t2_empty(a));
        add_packed<CS>(o, v2);
      }
      RB_TICK(1);
      if (ri >= RB_ROWS || m >= P.R) continue;
      {
        const float4 *bp = reinterpret_cast<const float4 *>(bias_s + NC + n0);
#pragma unroll
        for (int k = 0; k < CS / 4; ++k) {
          const float4 b = bp[k];
          add_packed<4>(o + 4 * k, &b.x);
        }
      }
      if (P.residual != nullptr && real) {
        const int lo0 = P.n_total >> 3;
#pragma unroll
        for (int g = 0; g < CS / 8; ++g) {
          const int pl = (n0 >> 3) + g;
          const uint4 rh = __ldg(P.residual + ((long long)pl * P.plane_rows + m));
          const uint4 rl = __ldg(P.residual + ((long long)(lo0 + pl) * P.plane_rows + m));
          float rr[8];
          join8(rh, rl, rr);
          add_packed<8>(o + g * 8, rr);
        }
      }
      elu_packed<CS>(o);
      store_planar<CS>(P, o, n0, m, c, r, t, real);
      RB_TICK(2);
    }
    if (pf) { for (int k = 0; k < 4; ++k) P.prof[24 + k] = pa[k]; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 17) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(8 * NC) : "memory");
  }
}

template <int NC>
int launch_pconv_rb_t(const KParams &P0, cudaStream_t st) {
  using C = Cfg<NC>;
  KParams P = P0;
  P.kps = (P.skip.kchunks % 2 == 0) ? 2 : 1;     // two 16-channel slices per stage when every source allows it: half the
  for (int i = 0; i < P.nsrc; ++i)               // barrier waits and commits on the MMA thread, which paces this kernel
    if (P.src[i].kchunks % 2) P.kps = 1;
  P.stage_bytes = P.kps * A_BYTES;               // activations only: the weights are resident
  int slices1 = 0;
  for (int i = 0; i < P.nsrc; ++i) slices1 += P.src[i].kchunks * P.src[i].taps;
  P.w1_bytes = slices1 * C::B_TAP;
  P.w2_bytes = ((NC / 16) * 3 + P.skip.kchunks) * C::B_TAP;
  const int u_bytes = 2 * 2 * (NC / 8) * U_PLANE;
  const int fixed = P.w1_bytes + P.w2_bytes + u_bytes + 8 * (2 * MAX_STAGE + 13) + 24 + 2 * NC * 4;
  int ns = (226 * 1024 - fixed) / P.stage_bytes;
  ns = ns > MAX_STAGE ? MAX_STAGE : ns;
  WM_CHECK_ARG(ns >= 3, "pconv_rb: %d bytes of resident weights leave no room for the activation stages", P.w1_bytes + P.w2_bytes);
  P.nstage = ns;
  P.u_off = ns * P.stage_bytes + P.w1_bytes + P.w2_bytes;
  const int smem_bytes = P.u_off + u_bytes + 8 * (2 * MAX_STAGE + 13) + 24 + 2 * NC * 4;
  static bool attr_set = false;
  if (!attr_set) {
    WM_CHECK_CUDA(cudaFuncSetAttribute(pconv_rb_kernel<NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  const long long ntiles = (P.R + RB_ROWS - 1) / RB_ROWS;
  WM_CHECK_ARG(P.R < (1LL << 31), "pconv_rb: %lld rows exceed the 32-bit tile index", P.R);
  const int grid = (int)(ntiles < sm_count() ? ntiles : sm_count());
  pconv_rb_kernel<NC><<<grid, 576, smem_bytes, st>>>(P);
  WM_CHECK_LAUNCH("pconv_rb");
  return 0;
}

// fp32 Wd[chunk][slice][16][NC] -> bf16 [chunk][slice][k8 = 0,1][n = 0..2NC-1][8]; n < NC: hi of column n, else lo of n - NC
__global__ void pconv_pack_kernel(const float *__restrict__ wd, __nv_bfloat16 *__restrict__ img, long long nslices, int nc) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long per = 32LL * nc;   // elements of one slice: 2 * (2 nc) * 8
  if (e >= nslices * per) return;
  const long long sl = e / per;
  const int w = (int)(e - sl * per);
  const int i = w & 7, n2 = (w >> 3) % (2 * nc), k8 = w / (16 * nc);
  const float v = wd[(sl * 16 + k8 * 8 + i) * nc + (n2 % nc)];
  const __nv_bfloat16 hi = __float2bfloat16_rn(v);
  img[e] = n2 < nc ? hi : __float2bfloat16_rn(v - __bfloat162float(hi));
}

// Conv1d(1, Cout, K, padding K/2) on the waveform, written planar and split into `split` phases
// (py/main14b_2.py:123,190 init_conv).  One thread per row of the phase geometry (T / split steps per clip).
__global__ void __launch_bounds__(256)
    pconv_in_kernel(const float *__restrict__ s, const float *__restrict__ w, const float *__restrict__ bias,
                    uint4 *__restrict__ y, int B, int T, int cout, int K, int split, long long plane_rows,
                    long long phase_rows) {
  extern __shared__ __align__(16) float ws[];   // [K][cout] then bias[cout]
  for (int i = threadIdx.x; i < K * cout; i += blockDim.x) ws[i] = w[(i % cout) * K + i / cout];
  for (int i = threadIdx.x; i < cout; i += blockDim.x) ws[K * cout + i] = bias[i];
  __syncthreads();
  const int Tq = T / split, Tp = Tq + GAP;
  const long long R = (long long)B * Tp + GAP;
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int phs = blockIdx.y;
  if (m >= R) return;
  const int c = (int)(m / Tp), r = (int)(m - (long long)c * Tp), qq = r - GAP;
  uint4 *yp = y + (long long)phs * phase_rows;
  const int lo0 = cout >> 3;
  if (c >= B || qq < 0) {
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (int pl = 0; pl < 2 * lo0; ++pl) yp[(long long)pl * plane_rows + m] = z;
    return;
  }
  const int t = qq * split + phs, P2 = K / 2;
  float sv[7];
#pragma unroll
  for (int k = 0; k < 7; ++k) {
    const int tt = t + k - P2;
    sv[k] = (k < K && tt >= 0 && tt < T) ? __ldg(s + (long long)c * T + tt) : 0.0f;
  }
  for (int g = 0; g < lo0; ++g) {
    float v[8];
    {
      const float4 b0 = *reinterpret_cast<const float4 *>(ws + K * cout + g * 8);
      const float4 b1 = *reinterpret_cast<const float4 *>(ws + K * cout + g * 8 + 4);
      v[0] = b0.x; v[1] = b0.y; v[2] = b0.z; v[3] = b0.w; v[4] = b1.x; v[5] = b1.y; v[6] = b1.z; v[7] = b1.w;
    }
#pragma unroll
    for (int k = 0; k < 7; ++k) {
      if (k < K) {
        const float4 w0 = *reinterpret_cast<const float4 *>(ws + k * cout + g * 8);
        const float4 w1 = *reinterpret_cast<const float4 *>(ws + k * cout + g * 8 + 4);
        v[0] = fmaf(sv[k], w0.x, v[0]); v[1] = fmaf(sv[k], w0.y, v[1]); v[2] = fmaf(sv[k], w0.z, v[2]);
        v[3] = fmaf(sv[k], w0.w, v[3]); v[4] = fmaf(sv[k], w1.x, v[4]); v[5] = fmaf(sv[k], w1.y, v[5]);
        v[6] = fmaf(sv[k], w1.z, v[6]); v[7] = fmaf(sv[k], w1.w, v[7]);
      }
    }
    uint4 hi, lo;
    split8(v, hi, lo);
    yp[(long long)g * plane_rows + m] = hi;
    yp[(long long)(lo0 + g) * plane_rows + m] = lo;
  }
}

// fp32 channels-first x[b][C][T] -> planar (gap rows zeroed)
__global__ void __launch_bounds__(256)
    pconv_to_planar_kernel(const float *__restrict__ x, uint4 *__restrict__ y, int B, int C, int T, long long plane_rows) {
  const int Tp = T + GAP;
  const long long R = (long long)B * Tp + GAP;
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int g = blockIdx.y, lo0 = C >> 3;
  if (m >= R) return;
  const int c = (int)(m / Tp), t = (int)(m - (long long)c * Tp) - GAP;
  uint4 hi = make_uint4(0, 0, 0, 0), lo = hi;
  if (c < B && t >= 0) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __ldg(x + ((long long)c * C + g * 8 + i) * T + t);
    split8(v, hi, lo);
  }
  y[(long long)g * plane_rows + m] = hi;
  y[(long long)(lo0 + g) * plane_rows + m] = lo;
}

// planar -> fp32 channels-first y[b][C][Tout] (first Tout steps)
__global__ void __launch_bounds__(256)
    pconv_from_planar_kernel(const uint4 *__restrict__ x, float *__restrict__ y, int B, int C, int T, int Tout,
                             long long plane_rows) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x, g = blockIdx.y, c = blockIdx.z, lo0 = C >> 3;
  if (t >= Tout) return;
  const long long m = (long long)c * (T + GAP) + GAP + t;
  const uint4 hi = __ldg(x + (long long)g * plane_rows + m), lo = __ldg(x + (long long)(lo0 + g) * plane_rows + m);
  float v[8];
  join8(hi, lo, v);
#pragma unroll
  for (int i = 0; i < 8; ++i) y[((long long)c * C + g * 8 + i) * Tout + t] = v[i];
}

}  // namespace

}  // namespace wm

using namespace wm;

extern "C" {

long long wm_pconv_plane_rows(int B, int T) {
  const long long R = (long long)B * (T + WM_PC_GAP) + WM_PC_GAP;
  return (R + 7) / 8 * 8 + 144;
}

size_t wm_pconv_desc_bytes(void) { return sizeof(wm_pconv); }

size_t wm_pconv_weight_bytes(long long nslices, int nc) { return (size_t)nslices * 64u * (size_t)nc; }

int wm_pconv_pack(const float *wd, void *img, long long nslices, int nc, void *stream) {
  if (int rc = require_device()) return rc;
  WM_CHECK_ARG(wd && img && nslices > 0, "pconv_pack: null pointer or no slices");
  WM_CHECK_ARG(nc == 16 || nc == 32 || nc == 64 || nc == 128, "pconv_pack: nc must be 16, 32, 64 or 128 (got %d)", nc);
  const long long n = nslices * 32 * nc;
  pconv_pack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(wd, reinterpret_cast<__nv_bfloat16 *>(img),
                                                                               nslices, nc);
  WM_CHECK_LAUNCH("pconv_pack");
  return 0;
}

int wm_pconv_fwd(const wm_pconv *d, void *stream) {
  if (int rc = require_device()) return rc;
  WM_CHECK_ARG(d != nullptr, "pconv: null descriptor");
  WM_CHECK_ARG(d->B >= 0 && d->T >= 0, "pconv: negative size");
  if (d->B == 0) return 0;
  WM_CHECK_ARG(d->nsrc >= 1 && d->nsrc <= 3, "pconv: 1..3 sources (got %d)", d->nsrc);
  WM_CHECK_ARG(d->nc == 16 || d->nc == 32 || d->nc == 64 || d->nc == 128, "pconv: nc must be 16, 32, 64 or 128 (got %d)",
               d->nc);
  WM_CHECK_ARG(d->n_total > 0 && d->n_total % d->nc == 0 && d->n_total / d->nc <= 64,
               "pconv: n_total %d must be a multiple of nc %d with at most 64 chunks", d->n_total, d->nc);
  WM_CHECK_ARG(d->w && d->bias && d->y, "pconv: null pointer");
  WM_CHECK_ARG(d->mode == WM_PC_OUT_PLANAR || d->mode == WM_PC_OUT_CONVT || d->mode == WM_PC_OUT_FP32,
               "pconv: unknown output mode %d", d->mode);
  KParams P{};
  long long slices = 0;
  for (int i = 0; i < d->nsrc; ++i) {
    const wm_pconv_src &s = d->src[i];
    WM_CHECK_ARG(s.base != nullptr, "pconv: source %d is null", i);
    WM_CHECK_ARG(s.cin >= 16 && s.cin % 16 == 0, "pconv: source %d has %d channels (multiple of 16 required)", i, s.cin);
    WM_CHECK_ARG(s.taps >= 1 && s.taps <= 7, "pconv: source %d has %d taps (1..7)", i, s.taps);
    WM_CHECK_ARG(s.row_off >= -WM_PC_GAP && s.row_off + s.taps - 1 <= WM_PC_GAP,
                 "pconv: source %d reaches beyond the %d gap rows", i, WM_PC_GAP);
    P.src[i].base = reinterpret_cast<const uint4 *>(s.base);
    P.src[i].kchunks = s.cin / 16;
    P.src[i].lo_plane = s.cin / 8;
    P.src[i].row_off = s.row_off;
    P.src[i].taps = s.taps;
    slices += (long long)(s.cin / 16) * s.taps;
  }
  P.nsrc = d->nsrc;
  P.plane_rows = d->plane_rows;
  P.B = d->B; P.T = d->T; P.Tp = d->T + WM_PC_GAP;
  P.R = (long long)d->B * P.Tp + WM_PC_GAP;
  WM_CHECK_ARG(d->plane_rows >= wm_pconv_plane_rows(d->B, d->T), "pconv: plane_rows %lld is below wm_pconv_plane_rows",
               d->plane_rows);
  P.w = reinterpret_cast<const uint8_t *>(d->w);
  P.w_chunk_bytes = slices * 64 * d->nc;
  P.bias = d->bias;
  P.nch = d->n_total / d->nc;
  P.n_total = d->n_total;
  P.elu = d->elu;
  P.residual = reinterpret_cast<const uint4 *>(d->residual);
  P.mode = d->mode;
  P.y = d->y;
  P.out_plane_rows = d->out_plane_rows;
  P.out_split = d->out_split < 1 ? 1 : d->out_split;
  P.out_phase_rows = d->out_phase_rows;
  P.ct_stride = d->ct_stride; P.ct_pad = d->ct_pad; P.ct_cout = d->ct_cout; P.out_T = d->out_T;
  P.ct_pk = d->ct_interleave < 1 ? 1 : d->ct_interleave;
  P.dbg = (get_debug_opts() >> 8) & 3;
  for (int i = 0; i < 64; ++i) P.chunk_off[i] = d->chunk_off[i];
  if (d->mode == WM_PC_OUT_PLANAR) {
    WM_CHECK_ARG(d->n_total % 8 == 0, "pconv: planar output needs a multiple of 8 channels");
    if (P.out_split > 1) {
      WM_CHECK_ARG(d->T % P.out_split == 0, "pconv: T %d is not a multiple of the output split %d", d->T, P.out_split);
      P.out_Tp = d->T / P.out_split + WM_PC_GAP;
      WM_CHECK_ARG(d->out_plane_rows >= wm_pconv_plane_rows(d->B, d->T / P.out_split), "pconv: out_plane_rows too small");
    } else {
      WM_CHECK_ARG(d->out_plane_rows >= wm_pconv_plane_rows(d->B, d->T), "pconv: out_plane_rows too small");
    }
  } else if (d->mode == WM_PC_OUT_CONVT) {
    WM_CHECK_ARG(d->ct_stride >= 1 && d->ct_cout >= 8 && d->ct_cout % 8 == 0 && d->n_total == d->ct_stride * d->ct_cout,
                 "pconv: transposed output needs n_total == stride * cout, cout a multiple of 8");
    WM_CHECK_ARG(d->out_T >= d->ct_stride * d->T - d->ct_stride && d->out_T <= d->ct_stride * (d->T + 1),
                 "pconv: transposed output length %d does not fit stride %d x %d rows", d->out_T, d->ct_stride, d->T);
    WM_CHECK_ARG(d->residual == nullptr, "pconv: no residual on the transposed output");
    WM_CHECK_ARG(d->ct_stride % P.ct_pk == 0, "pconv: ct_interleave %d does not divide the stride %d", P.ct_pk, d->ct_stride);
    P.out_Tp = d->out_T + WM_PC_GAP;
    WM_CHECK_ARG(d->out_plane_rows >= wm_pconv_plane_rows(d->B, d->out_T), "pconv: out_plane_rows too small");
  } else {
    WM_CHECK_ARG(d->ct_cout >= 1 && d->ct_cout <= d->n_total && d->out_T >= 0 && d->out_T <= d->T,
                 "pconv: fp32 output needs 1 <= channels <= n_total and out_T <= T");
  }
  cudaStream_t st = as_stream(stream);
  if (d->fused) {
    WM_CHECK_ARG(d->mode == WM_PC_OUT_PLANAR && d->n_total == d->nc && d->nc <= 64,
                 "pconv: the fused residual block needs planar output and one chunk of at most 64 columns");
    WM_CHECK_ARG(d->w2 && d->bias2, "pconv: the fused residual block needs w2 and bias2");
    WM_CHECK_ARG(d->chunk_off[0] == 0, "pconv: no chunk offset in the fused residual block");
    P.w2 = reinterpret_cast<const uint8_t *>(d->w2);
    P.bias2 = d->bias2;
    P.prof = get_profile_buffer();
    if (d->skip.base != nullptr) {
      WM_CHECK_ARG(d->skip.cin >= 16 && d->skip.cin % 16 == 0 && d->skip.taps == 1 && d->skip.row_off == 0,
                   "pconv: the skip source is one tap at row offset 0 over a multiple of 16 channels");
      P.skip.base = reinterpret_cast<const uint4 *>(d->skip.base);
      P.skip.kchunks = d->skip.cin / 16;
      P.skip.lo_plane = d->skip.cin / 8;
      P.skip.row_off = 0;
      P.skip.taps = 1;
    }
    switch (d->nc) {
      case 16: return launch_pconv_rb_t<16>(P, st);
      case 32: return launch_pconv_rb_t<32>(P, st);
      default: return launch_pconv_rb_t<64>(P, st);
    }
  }
  switch (d->nc) {
    case 16: return launch_pconv_t<16>(P, st);
    case 32: return launch_pconv_t<32>(P, st);
    case 64: return launch_pconv_t<64>(P, st);
    default: return launch_pconv_t<128>(P, st);
  }
}

int wm_pconv_in_fwd(const float *s, const float *w, const float *bias, void *y, int B, int T, int cout, int K, int split,
                    long long plane_rows, void *stream) {
  if (int rc = require_device()) return rc;
  WM_CHECK_ARG(B >= 0 && T >= 0, "pconv_in: negative size");
  if (B == 0) return 0;
  WM_CHECK_ARG(s && w && bias && y, "pconv_in: null pointer");
  WM_CHECK_ARG(cout >= 8 && cout % 8 == 0 && cout <= 128, "pconv_in: cout %d must be a multiple of 8, at most 128", cout);
  WM_CHECK_ARG(K >= 1 && K <= 7 && (K & 1), "pconv_in: K %d must be odd, at most 7", K);
  WM_CHECK_ARG(split >= 1 && T % split == 0, "pconv_in: T %d is not a multiple of split %d", T, split);
  WM_CHECK_ARG(plane_rows >= wm_pconv_plane_rows(B, T / split), "pconv_in: plane_rows too small");
  const long long R = (long long)B * (T / split + WM_PC_GAP) + WM_PC_GAP;
  dim3 grid((unsigned)((R + 255) / 256), split);
  pconv_in_kernel<<<grid, 256, (K + 1) * cout * sizeof(float), as_stream(stream)>>>(
      s, w, bias, reinterpret_cast<uint4 *>(y), B, T, cout, K, split, plane_rows, 2LL * (cout / 8) * plane_rows);
  WM_CHECK_LAUNCH("pconv_in");
  return 0;
}

int wm_pconv_to_planar(const float *x, void *y, int B, int C, int T, long long plane_rows, void *stream) {
  if (int rc = require_device()) return rc;
  WM_CHECK_ARG(B >= 0 && T >= 0 && C >= 8 && C % 8 == 0, "pconv_to_planar: bad size");
  if (B == 0) return 0;
  WM_CHECK_ARG(x && y, "pconv_to_planar: null pointer");
  WM_CHECK_ARG(plane_rows >= wm_pconv_plane_rows(B, T), "pconv_to_planar: plane_rows too small");
  const long long R = (long long)B * (T + WM_PC_GAP) + WM_PC_GAP;
  dim3 grid((unsigned)((R + 255) / 256), C / 8);
  pconv_to_planar_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, reinterpret_cast<uint4 *>(y), B, C, T, plane_rows);
  WM_CHECK_LAUNCH("pconv_to_planar");
  return 0;
}

int wm_pconv_from_planar(const void *x, float *y, int B, int C, int T, int Tout, long long plane_rows, void *stream) {
  if (int rc = require_device()) return rc;
  WM_CHECK_ARG(B >= 0 && T >= 0 && C >= 8 && C % 8 == 0 && Tout >= 0 && Tout <= T, "pconv_from_planar: bad size");
  if (B == 0 || Tout == 0) return 0;
  WM_CHECK_ARG(x && y, "pconv_from_planar: null pointer");
  WM_CHECK_ARG(B <= 65535, "pconv_from_planar: at most 65535 clips per call");
  dim3 grid((Tout + 255) / 256, C / 8, B);
  pconv_from_planar_kernel<<<grid, 256, 0, as_stream(stream)>>>(reinterpret_cast<const uint4 *>(x), y, B, C, T, Tout,
                                                               plane_rows);
  WM_CHECK_LAUNCH("pconv_from_planar");
  return 0;
}

}  // extern "C"
