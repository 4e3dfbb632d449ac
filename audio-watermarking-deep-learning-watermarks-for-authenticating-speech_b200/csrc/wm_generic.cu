// Generic fp32 operators for the configurable residual stack of main14b_2 (py/main14b_2.py:86-224, BASELINE
// config 3): strided Conv1d and strided ConvTranspose1d of any channel counts with fused bias / per-clip
// channel add / residual add / ELU epilogues, and the small 2-layer LSTM(32) bottleneck.  Tensors are
// channels-first fp32 x[b][c][t] exactly as the reference holds them; weights are the reference's own
// parameter tensors (Conv1d (co,ci,k), ConvTranspose1d (ci,co,k)), no packing.
// First correct CUDA path for this model family: CUDA-core FMA with shared-memory tiling (the tcgen05
// kernels of the main16 path are specialised to 64 channels, stride 1).
#include <stdlib.h>

#include "wm_common.h"

namespace wm {

namespace {

constexpr int GT = 64;        // output time steps per block
constexpr int GCO = 64;       // output channels per block
constexpr int GCI = 8;        // input channels per shared-memory stage
constexpr int GTHREADS = 256; // 16 (time) x 16 (channel) threads, 4 x 4 outputs each
constexpr int GMAXK = 16;     // taps
constexpr int GMAXS = 8;      // stride

__device__ __forceinline__ float elu1(float v) { return v > 0.0f ? v : expm1f(v); }

// y[b][co][t] = act( bias[co] + chan_add[b][co] + sum_ci sum_k w[co][ci][k] x[b][ci][t*stride + k - pad] + res[b][co][t] )
__global__ void __launch_bounds__(GTHREADS)
    conv1d_generic_kernel(const float *__restrict__ x, const float *__restrict__ w, const float *__restrict__ bias,
                          const float *__restrict__ chan_add, const float *__restrict__ res, float *__restrict__ y,
                          int Cin, int Tin, int Cout, int Tout, int K, int stride, int pad, int act, int shuffle,
                          int Tstore) {
  extern __shared__ float sm[];
  const int span = (GT - 1) * stride + K;            // input samples one stage needs per channel
  float *xs = sm;                                    // [GCI][span]
  float *ws = sm + GCI * span;                       // [GCI][K][GCO]  (co contiguous)
  const int b = blockIdx.z, co0 = blockIdx.y * GCO, t0 = blockIdx.x * GT;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;     // time group, channel group
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
  const float *xb = x + (size_t)b * Cin * Tin;
  const int in0 = t0 * stride - pad;
  for (int c0 = 0; c0 < Cin; c0 += GCI) {
    __syncthreads();
    for (int e = threadIdx.x; e < GCI * span; e += GTHREADS) {
      const int ci = e / span, i = e - ci * span, ti = in0 + i;
      xs[e] = (c0 + ci < Cin && ti >= 0 && ti < Tin) ? xb[(size_t)(c0 + ci) * Tin + ti] : 0.0f;
    }
    for (int e = threadIdx.x; e < GCI * K * GCO; e += GTHREADS) {
      const int co = e % GCO, r = e / GCO, k = r % K, ci = r / K;
      ws[e] = (c0 + ci < Cin && co0 + co < Cout) ? w[((size_t)(co0 + co) * Cin + c0 + ci) * K + k] : 0.0f;
    }
    __syncthreads();
    for (int ci = 0; ci < GCI; ++ci) {
      for (int k = 0; k < K; ++k) {
        const float4 wv = *reinterpret_cast<const float4 *>(&ws[(ci * K + k) * GCO + ty * 4]);
        const float *xp = &xs[ci * span + k];
        float xv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) xv[j] = xp[(tx + 16 * j) * stride];
        const float wq[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(wq[i], xv[j], acc[i][j]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = co0 + ty * 4 + i;
    if (co >= Cout) continue;
    const float add = bias[co] + (chan_add ? chan_add[(size_t)b * Cout + co] : 0.0f);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int t = t0 + tx + 16 * j;
      if (t >= Tout) continue;
      float v = acc[i][j] + add;
      if (shuffle > 1) {   // channel co = (c, r) is phase r of output channel c: y[b][c][t * shuffle + r]
        const int ts = t * shuffle + co % shuffle;
        if (ts < Tstore) y[((size_t)b * (Cout / shuffle) + co / shuffle) * Tstore + ts] = v;
        continue;
      }
      const size_t o = ((size_t)b * Cout + co) * Tout + t;
      if (res) v += res[o];
      y[o] = act ? elu1(v) : v;
    }
  }
}

// ---- the same convolution as a register-tiled implicit GEMM ---------------------------------------------------------
// M = output channels, N = (clip, output time) flattened, Kdim = Cin * K.  A block of 128 threads owns COG*8 output
// channels x (128/COG)*8 positions; a thread owns 8 channels x 8 positions.  Per stage of FCI input
// channels the block builds in shared memory  xs[ci][k][n] = x[b(n)][c0+ci][t(n)*stride + k - pad]  (an im2col slice:
// every tap gets its own aligned row, so strided convolutions read only the phases they use and all inner-loop loads
// are LDS.128) and  ws[ci][k][co]  (co contiguous);  per (ci, k): 2 + 2 LDS.128 feed 64 FMAs.
__device__ __forceinline__ unsigned long long pack2(float a, float b) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float &a, float &b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ unsigned long long fma2x(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

__device__ __forceinline__ void cp_async4_zfill(void *smem, const void *gmem, int src_bytes) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(sa), "l"(gmem), "r"(src_bytes));
}

template <int COG, int FCI>
__global__ void __launch_bounds__(128, 4)
    conv1d_tiled_kernel(const float *__restrict__ x, const float *__restrict__ w, const float *__restrict__ bias,
                        const float *__restrict__ chan_add, const float *__restrict__ res, float *__restrict__ y,
                        int Cin, int Tin, int Cout, int Tout, int K, int stride, int pad, int act, int shuffle, int Tstore,
                        long long Ntot) {
  constexpr int CO_T = COG * 8, TG = 128 / COG, N_T = TG * 8;
  constexpr int WLD = CO_T + 4;   // pitch of a weight row: consecutive (ci,k) rows start 4 banks apart (the staging writes
                                  // walk down a column)
  extern __shared__ __align__(16) float sm[];          // two stages of { xs [FCI][K][N_T], ws [FCI][K][WLD] }
  const int co0 = blockIdx.y * CO_T;
  const long long n0 = (long long)blockIdx.x * N_T;
  const int cg = threadIdx.x % COG, tg = threadIdx.x / COG;
  unsigned long long acc2[8][4];     // [channel][position pair], two fp32 each
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc2[i][j] = 0ull;
  // staging ownership: thread -> positions n0 + threadIdx.x + 128 q; which of the K taps fall inside [0, Tin) is a
  // per-position bit mask computed once
  constexpr int NQ = N_T / 128;
  const float *xpos[NQ];      // &x[b][0][t*stride - pad]
  unsigned kmask[NQ];
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    const unsigned n = (unsigned)n0 + threadIdx.x + 128 * q;
    xpos[q] = x;
    kmask[q] = 0u;
    if (n < (unsigned)Ntot) {
      const int b = (int)(n / (unsigned)Tout), t = (int)(n - (unsigned)b * (unsigned)Tout), ti0 = t * stride - pad;
      xpos[q] = x + (size_t)b * Cin * Tin + ti0;
      for (int k = 0; k < K; ++k)
        if (ti0 + k >= 0 && ti0 + k < Tin) kmask[q] |= 1u << k;
    }
  }
  const int wrun = FCI * K;                         // contiguous weights per output channel and stage
  const int stage_floats = FCI * K * (N_T + WLD);
  const int w_dq = 128 / wrun, w_dr = 128 % wrun;   // e += 128  <=>  (co, r) += (w_dq, w_dr) with carry
  const int w_co_first = threadIdx.x / wrun, w_r_first = threadIdx.x % wrun;
  // 4-byte cp.async (zero-filled when out of range): the ~36 loads a thread issues per stage are all in flight at
  // once, and the next stage is fetched while this one is multiplied
  auto stage = [&](int c0, int buf) {
    float *xs = sm + buf * stage_floats, *ws = xs + FCI * K * N_T;
    const int nci = min(FCI, Cin - c0);
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      float *dst = xs + threadIdx.x + 128 * q;
      const float *src = xpos[q] + (size_t)c0 * Tin;
      for (int ci = 0; ci < nci; ++ci, src += Tin) {
        for (int k = 0; k < K; ++k, dst += N_T) {
          const bool ok = (kmask[q] >> k) & 1u;
          cp_async4_zfill(dst, ok ? src + k : x, ok ? 4 : 0);
        }
      }
    }
    const int nrk = nci * K;
    int co = w_co_first, r = w_r_first;
    for (int e = threadIdx.x; e < CO_T * wrun; e += 128) {
      const bool ok = co0 + co < Cout && r < nrk;
      cp_async4_zfill(&ws[r * WLD + co], ok ? w + ((size_t)(co0 + co) * Cin + c0) * K + r : w, ok ? 4 : 0);
      co += w_dq;
      r += w_dr;
      if (r >= wrun) { r -= wrun; ++co; }
    }
    asm volatile("cp.async.commit_group;\n");
  };
  stage(0, 0);
  int it = 0;
  for (int c0 = 0; c0 < Cin; c0 += FCI, ++it) {
    const int buf = it & 1;
    const int nci = min(FCI, Cin - c0);
    if (c0 + FCI < Cin) {
      stage(c0 + FCI, buf ^ 1);
      asm volatile("cp.async.wait_group 1;\n");
    } else {
      asm volatile("cp.async.wait_group 0;\n");
    }
    __syncthreads();
    const float *xs = sm + buf * stage_floats, *ws = xs + FCI * K * N_T;
    const int nr = nci * K;
#pragma unroll 2
    for (int r = 0; r < nr; ++r) {
      // a thread's 8 channels / positions are two runs of 4, half a tile apart: consecutive lanes read consecutive
      // 16-byte words (no bank conflicts)
      const float4 w0 = *reinterpret_cast<const float4 *>(&ws[r * WLD + cg * 4]);
      const float4 w1 = *reinterpret_cast<const float4 *>(&ws[r * WLD + CO_T / 2 + cg * 4]);
      // packed fp32 FMAs (fma.rn.f32x2): positions are read as 64-bit pairs, the weight is duplicated
      const ulonglong2 x0 = *reinterpret_cast<const ulonglong2 *>(&xs[r * N_T + tg * 4]);
      const ulonglong2 x1 = *reinterpret_cast<const ulonglong2 *>(&xs[r * N_T + N_T / 2 + tg * 4]);
      const float wq[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
      const unsigned long long xp[4] = {x0.x, x0.y, x1.x, x1.y};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const unsigned long long ww = pack2(wq[i], wq[i]);
#pragma unroll
        for (int j = 0; j < 4; ++j) acc2[i][j] = fma2x(ww, xp[j], acc2[i][j]);
      }
    }
    __syncthreads();   // the buffer is refilled by the stage() issued at the top of the next iteration
  }
  // epilogue: the thread's positions are two runs of 4 consecutive n; with Tout % 4 == 0 a run never straddles a clip
  // and is 16-byte aligned in y, so it is stored (and the residual read) as one float4.  32-bit index math: the
  // launcher guarantees B * Tout < 2^31.
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) unpack2(acc2[i][j], acc[i][2 * j], acc[i][2 * j + 1]);
  const unsigned Nt = (unsigned)Ntot, uT = (unsigned)Tout;
  const bool vec = shuffle <= 1 && (Tout & 3) == 0;
#pragma unroll
  for (int jr = 0; jr < 2; ++jr) {
    const unsigned nb = (unsigned)n0 + (jr ? N_T / 2 : 0) + tg * 4;
    if (nb >= Nt) continue;
    const unsigned b0 = nb / uT, t0 = nb - b0 * uT;
    // phase-interleaved (transposed-convolution) output with stride 2, 4 or 8: a run of 4 channels x 4 positions is
    // whole float4 pieces of the interleaved rows  y[b][co / s][t * s + co % s]
    if ((shuffle == 2 || shuffle == 4 || shuffle == 8) && (Tstore & 3) == 0 && t0 + 3 < uT &&
        (int)(t0 + 4) * shuffle <= Tstore) {
#pragma unroll
      for (int ir = 0; ir < 2; ++ir) {
        const int cob = co0 + (ir ? CO_T / 2 : 0) + cg * 4;          // first of 4 consecutive output rows
        if (cob + 3 >= Cout) {
          if (cob >= Cout) continue;
        }
        float v[4][4];                                                 // [row][position] with bias (and clip add)
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          const int co = cob + a;
          const float add = co < Cout ? bias[co] + (chan_add ? chan_add[(size_t)b0 * Cout + co] : 0.0f) : 0.0f;
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) v[a][jj] = acc[ir * 4 + a][jr * 4 + jj] + add;
        }
        const int C = Cout / shuffle;
        if (cob + 3 < Cout) {
          if (shuffle == 4) {
            float *dst = y + ((size_t)b0 * C + cob / 4) * Tstore + (size_t)t0 * 4;
#pragma unroll
            for (int jj = 0; jj < 4; ++jj)
              *reinterpret_cast<float4 *>(dst + 4 * jj) = make_float4(v[0][jj], v[1][jj], v[2][jj], v[3][jj]);
          } else if (shuffle == 8) {
            float *dst = y + ((size_t)b0 * C + cob / 8) * Tstore + (size_t)t0 * 8 + (cob & 7);
#pragma unroll
            for (int jj = 0; jj < 4; ++jj)
              *reinterpret_cast<float4 *>(dst + 8 * jj) = make_float4(v[0][jj], v[1][jj], v[2][jj], v[3][jj]);
          } else {
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
              float *dst = y + ((size_t)b0 * C + cob / 2 + cc) * Tstore + (size_t)t0 * 2;
              *reinterpret_cast<float4 *>(dst) = make_float4(v[2 * cc][0], v[2 * cc + 1][0], v[2 * cc][1], v[2 * cc + 1][1]);
              *reinterpret_cast<float4 *>(dst + 4) = make_float4(v[2 * cc][2], v[2 * cc + 1][2], v[2 * cc][3], v[2 * cc + 1][3]);
            }
          }
          continue;
        }
        for (int a = 0; a < 4 && cob + a < Cout; ++a)                  // ragged last rows: scalar
          for (int jj = 0; jj < 4; ++jj) {
            const int co = cob + a, ts = (int)(t0 + jj) * shuffle + co % shuffle;
            if (ts < Tstore) y[((size_t)b0 * C + co / shuffle) * Tstore + ts] = v[a][jj];
          }
      }
      continue;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int co = co0 + (i < 4 ? cg * 4 + i : CO_T / 2 + cg * 4 + i - 4);
      if (co >= Cout) continue;
      const float bv = bias[co];
      if (vec) {
        const float add = bv + (chan_add ? chan_add[(size_t)b0 * Cout + co] : 0.0f);
        const size_t o = ((size_t)b0 * Cout + co) * Tout + t0;
        float4 v = make_float4(acc[i][jr * 4] + add, acc[i][jr * 4 + 1] + add, acc[i][jr * 4 + 2] + add,
                               acc[i][jr * 4 + 3] + add);
        if (res) {
          const float4 r = *reinterpret_cast<const float4 *>(res + o);
          v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
        }
        if (act) { v.x = elu1(v.x); v.y = elu1(v.y); v.z = elu1(v.z); v.w = elu1(v.w); }
        *reinterpret_cast<float4 *>(y + o) = v;
        continue;
      }
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const unsigned n = nb + jj;
        if (n >= Nt) break;
        unsigned b = b0, t = t0 + jj;
        if (t >= uT) { b = n / uT; t = n - b * uT; }
        float v = acc[i][jr * 4 + jj] + bv + (chan_add ? chan_add[(size_t)b * Cout + co] : 0.0f);
        if (shuffle > 1) {   // channel co = (c, r) is phase r of output channel c: y[b][c][t * shuffle + r]
          const int ts = (int)t * shuffle + co % shuffle;
          if (ts < Tstore) y[((size_t)b * (Cout / shuffle) + co / shuffle) * Tstore + ts] = v;
          continue;
        }
        const size_t o = ((size_t)b * Cout + co) * Tout + t;
        if (res) v += res[o];
        y[o] = act ? elu1(v) : v;
      }
    }
  }
}

template <int COG, int FCI>
int launch_conv1d_tiled(const float *x, const float *w, const float *bias, const float *chan_add, const float *res,
                        float *y, int B, int Cin, int Tin, int Cout, int Tout, int K, int stride, int pad, int act,
                        int shuffle, int Tstore, cudaStream_t st) {
  constexpr int CO_T = COG * 8, N_T = (128 / COG) * 8;
  const size_t smem = 2 * (size_t)FCI * K * (N_T + CO_T + 4) * sizeof(float);
  static size_t attr = 0;
  if (smem > attr) {
    WM_CHECK_CUDA(cudaFuncSetAttribute(conv1d_tiled_kernel<COG, FCI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem));
    attr = smem;
  }
  const long long Ntot = (long long)B * Tout;
  dim3 grid((unsigned)((Ntot + N_T - 1) / N_T), (Cout + CO_T - 1) / CO_T);
  conv1d_tiled_kernel<COG, FCI><<<grid, 128, smem, st>>>(x, w, bias, chan_add, res, y, Cin, Tin, Cout, Tout, K, stride, pad,
                                                         act, shuffle, Tstore, Ntot);
  WM_CHECK_LAUNCH("conv1d_tiled");
  return 0;
}

// nn.ConvTranspose1d(Cin, Cout, K, stride, padding):  y[b][co][t] = bias[co] + sum_ci sum_k [ (t + pad - k) % stride == 0 ]
//   x[b][ci][(t + pad - k) / stride] w[ci][co][k]
__global__ void __launch_bounds__(GTHREADS)
    convtranspose1d_generic_kernel(const float *__restrict__ x, const float *__restrict__ w, const float *__restrict__ bias,
                                   float *__restrict__ y, int Cin, int Tin, int Cout, int Tout, int K, int stride, int pad) {
  extern __shared__ float sm[];
  // inputs that can reach outputs t0 .. t0+GT-1:  i in [ceil((t0 + pad - K + 1) / stride), floor((t0 + GT - 1 + pad) / stride)]
  const int b = blockIdx.z, co0 = blockIdx.y * GCO, t0 = blockIdx.x * GT;
  const int nspan = (GT + K - 2) / stride + 2;
  float *xs = sm;                                    // [GCI][nspan]
  float *ws = sm + GCI * nspan;                      // [GCI][K][GCO]
  int ilo = t0 + pad - K + 1;
  ilo = ilo >= 0 ? (ilo + stride - 1) / stride : -((-ilo) / stride);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
  const float *xb = x + (size_t)b * Cin * Tin;
  for (int c0 = 0; c0 < Cin; c0 += GCI) {
    __syncthreads();
    for (int e = threadIdx.x; e < GCI * nspan; e += GTHREADS) {
      const int ci = e / nspan, i = e - ci * nspan, ti = ilo + i;
      xs[e] = (c0 + ci < Cin && ti >= 0 && ti < Tin) ? xb[(size_t)(c0 + ci) * Tin + ti] : 0.0f;
    }
    for (int e = threadIdx.x; e < GCI * K * GCO; e += GTHREADS) {
      const int co = e % GCO, r = e / GCO, k = r % K, ci = r / K;
      ws[e] = (c0 + ci < Cin && co0 + co < Cout) ? w[((size_t)(c0 + ci) * Cout + co0 + co) * K + k] : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int t = t0 + tx + 16 * j;
      // taps k = (t + pad) % stride, + stride, ... < K
      for (int k = (t + pad) % stride; k < K; k += stride) {
        const int i = (t + pad - k) / stride - ilo;   // exact division; may fall outside [0, Tin): staged as zero
        if (i < 0 || i >= nspan) continue;
        for (int ci = 0; ci < GCI; ++ci) {
          const float xv = xs[ci * nspan + i];
          const float4 wv = *reinterpret_cast<const float4 *>(&ws[(ci * K + k) * GCO + ty * 4]);
          acc[0][j] = fmaf(wv.x, xv, acc[0][j]);
          acc[1][j] = fmaf(wv.y, xv, acc[1][j]);
          acc[2][j] = fmaf(wv.z, xv, acc[2][j]);
          acc[3][j] = fmaf(wv.w, xv, acc[3][j]);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = co0 + ty * 4 + i;
    if (co >= Cout) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int t = t0 + tx + 16 * j;
      if (t < Tout) y[((size_t)b * Cout + co) * Tout + t] = acc[i][j] + bias[co];
    }
  }
}

// nn.LSTM(H, H, num_layers = L, batch_first) on channels-first x[b][H][T] -> y[b][H][T] (top layer's hidden states),
// zero initial state, gate rows i,f,g,o.  One block per clip, 4H threads (thread = gate row); H <= 64, L <= 4.
// w_ih / w_hh: [L][4H][H], bias: [L][4H] = b_ih + b_hh.
__global__ void lstm_small_kernel(const float *__restrict__ x, const float *__restrict__ w_ih, const float *__restrict__ w_hh,
                                  const float *__restrict__ bias, float *__restrict__ y, int H, int T, int L) {
  extern __shared__ float sm[];
  float *h = sm;                 // [L][H]
  float *c = h + L * H;          // [L][H]
  float *g = c + L * H;          // [4H] gate pre-activations of the current layer
  float *xin = g + 4 * H;        // [H] input of the current layer
  const int b = blockIdx.x, r = threadIdx.x;
  for (int i = r; i < 2 * L * H; i += blockDim.x) h[i] = 0.0f;
  __syncthreads();
  for (int t = 0; t < T; ++t) {
    for (int l = 0; l < L; ++l) {
      if (r < H) xin[r] = l == 0 ? x[((size_t)b * H + r) * T + t] : h[(l - 1) * H + r];
      __syncthreads();
      const float *wi = w_ih + ((size_t)l * 4 * H + r) * H, *wh = w_hh + ((size_t)l * 4 * H + r) * H;
      float a = bias[l * 4 * H + r];
      for (int k = 0; k < H; ++k) a = fmaf(wi[k], xin[k], fmaf(wh[k], h[l * H + k], a));
      g[r] = a;
      __syncthreads();
      if (r < H) {
        const float ig = 1.0f / (1.0f + expf(-g[r])), fg = 1.0f / (1.0f + expf(-g[H + r]));
        const float gg = tanhf(g[2 * H + r]), og = 1.0f / (1.0f + expf(-g[3 * H + r]));
        const float cn = fg * c[l * H + r] + ig * gg;
        c[l * H + r] = cn;
        h[l * H + r] = og * tanhf(cn);
      }
      __syncthreads();
    }
    if (r < H) y[((size_t)b * H + r) * T + t] = h[(L - 1) * H + r];
  }
}

// w (Cin, Cout, 2s) of a ConvTranspose1d(stride s, padding p) -> w3 (Cout*s, Cin, 3), b3 (Cout*s):
// output t = m*s + r takes x[m + a_r] w[.., k_r] + x[m + a_r - 1] w[.., k_r + s], k_r = (r + p) % s, a_r = (r + p) / s
__global__ void convt_phase_weights_kernel(const float *__restrict__ w, const float *__restrict__ bias, float *__restrict__ w3,
                                           float *__restrict__ b3, int Cin, int Cout, int s, int p) {
  const long long n = (long long)Cout * s * Cin;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const int ci = (int)(e % Cin);
    const int cr = (int)(e / Cin), co = cr / s, r = cr % s;
    const int kr = (r + p) % s, ar = (r + p) / s;
    const float *src = w + ((size_t)ci * Cout + co) * 2 * s;
    float t3[3] = {0.0f, 0.0f, 0.0f};
    t3[ar + 1] = src[kr];          // window offset d = a_r      -> tap 1 + d
    t3[ar] = src[kr + s];          // window offset d = a_r - 1
    w3[e * 3] = t3[0]; w3[e * 3 + 1] = t3[1]; w3[e * 3 + 2] = t3[2];
    if (ci == 0) b3[cr] = bias[co];
  }
}

}  // namespace

int launch_convt_phase_weights(const float *w, const float *bias, float *w3, float *b3, int Cin, int Cout, int s, int p,
                               cudaStream_t st) {
  const long long n = (long long)Cout * s * Cin;
  const long long nb = (n + 255) / 256;
  convt_phase_weights_kernel<<<(int)(nb < 4096 ? nb : 4096), 256, 0, st>>>(w, bias, w3, b3, Cin, Cout, s, p);
  WM_CHECK_LAUNCH("convt_phase_weights");
  return 0;
}

// shuffle > 1 (with Tstore): the ConvTranspose1d-as-convolution form.  A stride-s transposed convolution with
// K = 2s taps is s interleaved 2-tap convolutions of the input (one per output phase), i.e. ONE stride-1 3-tap
// convolution with Cout*s output channels whose results are written phase-interleaved; `extra_out` computes that
// many more positions than the input has (the last phase-0 sample of an odd stride).
int launch_conv1d_generic(const float *x, const float *w, const float *bias, const float *chan_add, const float *res,
                          float *y, int B, int Cin, int Tin, int Cout, int K, int stride, int pad, int act,
                          cudaStream_t st, int shuffle, int Tstore, int extra_out) {
  const int Tout = (Tin + 2 * pad - K) / stride + 1 + extra_out;
  if (B == 0 || Tout <= 0) return 0;
  if (K > GMAXK || stride > GMAXS || K < 1 || stride < 1) {
    set_error("conv1d: kernel size %d / stride %d outside the supported range (<= %d / <= %d)", K, stride, GMAXK, GMAXS);
    return -1;
  }
  // register-tiled implicit GEMM: 64 channels x 128 positions per block, or 16 x 512 for the narrow heads
  if (getenv("WMB200_CONV1D_SIMPLE") == nullptr && (long long)B * Tout < (1LL << 31) - 1024) {
    if (Cout > 16 && Cout <= 32) {   // one 32-channel tile x 256 positions: the input slice is staged once, not twice
      if (K <= 8) return launch_conv1d_tiled<4, 4>(x, w, bias, chan_add, res, y, B, Cin, Tin, Cout, Tout, K, stride, pad, act, shuffle, Tstore, st);
      return launch_conv1d_tiled<4, 2>(x, w, bias, chan_add, res, y, B, Cin, Tin, Cout, Tout, K, stride, pad, act, shuffle, Tstore, st);
    }
    if (Cout > 16) {
      if (K <= 8) return launch_conv1d_tiled<8, 8>(x, w, bias, chan_add, res, y, B, Cin, Tin, Cout, Tout, K, stride, pad, act, shuffle, Tstore, st);
      return launch_conv1d_tiled<8, 4>(x, w, bias, chan_add, res, y, B, Cin, Tin, Cout, Tout, K, stride, pad, act, shuffle, Tstore, st);
    }
    if (K <= 8) return launch_conv1d_tiled<2, 4>(x, w, bias, chan_add, res, y, B, Cin, Tin, Cout, Tout, K, stride, pad, act, shuffle, Tstore, st);
    return launch_conv1d_tiled<2, 2>(x, w, bias, chan_add, res, y, B, Cin, Tin, Cout, Tout, K, stride, pad, act, shuffle, Tstore, st);
  }
  const size_t smem = (size_t)(GCI * ((GT - 1) * stride + K) + GCI * K * GCO) * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    WM_CHECK_CUDA(cudaFuncSetAttribute(conv1d_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    attr_set = true;
  }
  dim3 grid((Tout + GT - 1) / GT, (Cout + GCO - 1) / GCO, B);
  conv1d_generic_kernel<<<grid, GTHREADS, smem, st>>>(x, w, bias, chan_add, res, y, Cin, Tin, Cout, Tout, K, stride, pad, act,
                                                      shuffle, Tstore);
  WM_CHECK_LAUNCH("conv1d_generic");
  return 0;
}

int launch_convtranspose1d_generic(const float *x, const float *w, const float *bias, float *y, int B, int Cin, int Tin,
                                   int Cout, int K, int stride, int pad, cudaStream_t st) {
  const int Tout = (Tin - 1) * stride - 2 * pad + K;
  if (B == 0 || Tout <= 0) return 0;
  if (K > 2 * GMAXK || stride > GMAXS || K < 1 || stride < 1) {
    set_error("convtranspose1d: kernel size %d / stride %d outside the supported range", K, stride);
    return -1;
  }
  const size_t smem = (size_t)(GCI * ((GT + K - 2) / stride + 2) + GCI * K * GCO) * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    WM_CHECK_CUDA(cudaFuncSetAttribute(convtranspose1d_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    attr_set = true;
  }
  dim3 grid((Tout + GT - 1) / GT, (Cout + GCO - 1) / GCO, B);
  convtranspose1d_generic_kernel<<<grid, GTHREADS, smem, st>>>(x, w, bias, y, Cin, Tin, Cout, Tout, K, stride, pad);
  WM_CHECK_LAUNCH("convtranspose1d_generic");
  return 0;
}

int launch_lstm_small(const float *x, const float *w_ih, const float *w_hh, const float *bias, float *y, int B, int H,
                      int T, int L, cudaStream_t st) {
  if (B == 0 || T == 0) return 0;
  if (H < 1 || H > 64 || L < 1 || L > 4) {
    set_error("lstm_small: hidden size %d / layers %d outside the supported range (<= 64 / <= 4)", H, L);
    return -1;
  }
  {
    const int rc = launch_lstm_small_reg(x, w_ih, w_hh, bias, y, B, H, T, L, st);   // H = 32, L <= 2: weights in registers
    if (rc <= 0) return rc;
  }
  const size_t smem = (size_t)(2 * L * H + 4 * H + H) * sizeof(float);
  lstm_small_kernel<<<B, 4 * H, smem, st>>>(x, w_ih, w_hh, bias, y, H, T, L);
  WM_CHECK_LAUNCH("lstm_small");
  return 0;
}

}  // namespace wm
