// Training-mode building blocks of the main16 path (py/main16.py:238-278), fp32 on the CUDA cores, channels-last
// activations x[n][64] (n = b * T + t): batch-statistics BatchNorm forward / backward, convolution weight and data
// gradients, 1x1-head and BCE gradients, Adam — and the detector's complete training step assembled from them
// (forward in train mode, detection + message BCE, backward, Adam on one flat parameter buffer).  Every reduction
// goes through per-block partial sums added in a fixed order, so a step is deterministic.
// The generator's backward (LSTM through 16 000 steps, STFT losses) is not built yet: this file is the groundwork of
// BASELINE config 4 and is exercised against PyTorch autograd in tests/test_train.py.
#include <stdlib.h>

#include "wm_common.h"

namespace wm {

namespace {

constexpr int NT = 256;

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// ---- per-channel sums over rows: partial[blk][k][64] (double), k < NK ---------------------------------------
// MODE 0: {z, z^2}            MODE 1: {d, d * zhat} with d = dout * [act > 0], zhat = (z - mean) * rstd
template <int MODE>
__global__ void __launch_bounds__(NT)
    chan_sums_kernel(const float *__restrict__ a, const float *__restrict__ b, const float *__restrict__ act,
                     const float *__restrict__ mean, const float *__restrict__ rstd, long long N, int rows_per_block,
                     double *__restrict__ partial) {
  __shared__ double red[2][16][64];
  const int c4 = (threadIdx.x & 15) * 4, rl = threadIdx.x >> 4;
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  const long long r1 = r0 + rows_per_block < N ? r0 + rows_per_block : N;
  float s0[4] = {0, 0, 0, 0}, s1[4] = {0, 0, 0, 0};
  float mu[4] = {0, 0, 0, 0}, rs[4] = {1, 1, 1, 1};
  if (MODE == 1) {
#pragma unroll
    for (int k = 0; k < 4; ++k) { mu[k] = mean[c4 + k]; rs[k] = rstd[c4 + k]; }
  }
  for (long long r = r0 + rl; r < r1; r += 16) {
    const float4 av = *reinterpret_cast<const float4 *>(a + r * 64 + c4);
    const float v[4] = {av.x, av.y, av.z, av.w};
    if (MODE == 0) {
#pragma unroll
      for (int k = 0; k < 4; ++k) { s0[k] += v[k]; s1[k] = fmaf(v[k], v[k], s1[k]); }
    } else {
      const float4 zv = *reinterpret_cast<const float4 *>(b + r * 64 + c4);
      const float z[4] = {zv.x, zv.y, zv.z, zv.w};
      float m[4] = {1, 1, 1, 1};
      if (act) {
        const float4 ov = *reinterpret_cast<const float4 *>(act + r * 64 + c4);
        m[0] = ov.x > 0.f; m[1] = ov.y > 0.f; m[2] = ov.z > 0.f; m[3] = ov.w > 0.f;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float d = v[k] * m[k];
        s0[k] += d;
        s1[k] = fmaf(d, (z[k] - mu[k]) * rs[k], s1[k]);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) { red[0][rl][c4 + k] = s0[k]; red[1][rl][c4 + k] = s1[k]; }
  __syncthreads();
  if (threadIdx.x < 128) {
    const int which = threadIdx.x >> 6, c = threadIdx.x & 63;
    double t = 0.0;
    for (int i = 0; i < 16; ++i) t += red[which][i][c];
    partial[((size_t)blockIdx.x * 2 + which) * 64 + c] = t;
  }
}

// out[k][c] = sum_blk partial[blk][k][c]   (k < 2), fixed order: 8 interleaved segments per output, then a fixed tree
__global__ void __launch_bounds__(1024) sum2_kernel(const double *__restrict__ partial, int nblk, double *__restrict__ out) {
  __shared__ double red[8][128];
  const int i = threadIdx.x & 127, seg = threadIdx.x >> 7;
  double t = 0.0;
  for (int b = seg; b < nblk; b += 8) t += partial[(size_t)b * 128 + i];
  red[seg][i] = t;
  __syncthreads();
  if (seg == 0)
    out[i] = ((red[0][i] + red[1][i]) + (red[2][i] + red[3][i])) + ((red[4][i] + red[5][i]) + (red[6][i] + red[7][i]));
}

// batch statistics from {sum z, sum z^2}; running stats as nn.BatchNorm1d (momentum 0.1, unbiased variance)
__global__ void bn_finish_stats_kernel(const double *__restrict__ sums, long long N, float eps, float momentum,
                                       float *__restrict__ mean, float *__restrict__ rstd, float *__restrict__ run_mean,
                                       float *__restrict__ run_var) {
  const int c = threadIdx.x;
  const double m = sums[c] / (double)N;
  double var = sums[64 + c] / (double)N - m * m;
  if (var < 0) var = 0;
  mean[c] = (float)m;
  rstd[c] = (float)(1.0 / sqrt(var + (double)eps));
  if (run_mean) {
    run_mean[c] = (1.0f - momentum) * run_mean[c] + momentum * (float)m;
    const double unb = N > 1 ? var * (double)N / (double)(N - 1) : var;
    run_var[c] = (1.0f - momentum) * run_var[c] + momentum * (float)unb;
  }
}

// out = relu?( gamma * (z - mean) * rstd + beta + residual )
__global__ void bn_apply_kernel(const float *__restrict__ z, const float *__restrict__ mean, const float *__restrict__ rstd,
                                const float *__restrict__ gamma, const float *__restrict__ beta,
                                const float *__restrict__ residual, float *__restrict__ out, long long n4, int relu) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i & 15) * 4;
    const float4 v = reinterpret_cast<const float4 *>(z)[i];
    float o[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) o[k] = gamma[c + k] * (o[k] - mean[c + k]) * rstd[c + k] + beta[c + k];
    if (residual) {
      const float4 r = reinterpret_cast<const float4 *>(residual)[i];
      o[0] += r.x; o[1] += r.y; o[2] += r.z; o[3] += r.w;
    }
    if (relu) {
#pragma unroll
      for (int k = 0; k < 4; ++k) o[k] = fmaxf(o[k], 0.0f);
    }
    reinterpret_cast<float4 *>(out)[i] = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// d = dout * [act > 0];  dz = gamma * rstd * (d - S1/N - zhat * S2/N);  dres (nullable) = d;  sums = {S1, S2}
__global__ void bn_bwd_apply_kernel(const float *__restrict__ dout, const float *__restrict__ act, const float *__restrict__ z,
                                    const float *__restrict__ mean, const float *__restrict__ rstd,
                                    const float *__restrict__ gamma, const double *__restrict__ sums, long long N,
                                    float *__restrict__ dz, float *__restrict__ dres, long long n4) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i & 15) * 4;
    const float4 dv = reinterpret_cast<const float4 *>(dout)[i], zv = reinterpret_cast<const float4 *>(z)[i];
    float d[4] = {dv.x, dv.y, dv.z, dv.w};
    const float zz[4] = {zv.x, zv.y, zv.z, zv.w};
    if (act) {
      const float4 ov = reinterpret_cast<const float4 *>(act)[i];
      d[0] *= ov.x > 0.f; d[1] *= ov.y > 0.f; d[2] *= ov.z > 0.f; d[3] *= ov.w > 0.f;
    }
    float o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float zh = (zz[k] - mean[c + k]) * rstd[c + k];
      const float m1 = (float)(sums[c + k] / (double)N), m2 = (float)(sums[64 + c + k] / (double)N);
      o[k] = gamma[c + k] * rstd[c + k] * (d[k] - m1 - zh * m2);
    }
    reinterpret_cast<float4 *>(dz)[i] = make_float4(o[0], o[1], o[2], o[3]);
    if (dres) reinterpret_cast<float4 *>(dres)[i] = make_float4(d[0], d[1], d[2], d[3]);
  }
}

// dgamma = S2, dbeta = S1
__global__ void bn_param_grads_kernel(const double *__restrict__ sums, float *__restrict__ dgamma, float *__restrict__ dbeta) {
  const int c = threadIdx.x;
  dbeta[c] = (float)sums[c];
  dgamma[c] = (float)sums[64 + c];
}

// w[j][ci][co] -> wt[K-1-j][co][ci]  (data-gradient weights of a stride-1 'same' convolution)
__global__ void transpose_flip_kernel(const float *__restrict__ w, float *__restrict__ wt, int K) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= K * 4096) return;
  const int co = e & 63, ci = (e >> 6) & 63, j = e >> 12;
  wt[((K - 1 - j) * 64 + co) * 64 + ci] = w[e];
}

// weight gradient of y[t][co] = sum_j sum_ci w[j][ci][co] x[t + j - P][ci], one tap per blockIdx.z:
// partial[blk][j][ci][co] = sum over the block's rows of x[t + j - P][ci] * dz[t][co];  bias partial on j == bias_tap.
// 64 rows at a time are staged in shared memory; the four 64-thread groups take 16 rows each and every thread owns
// an 8 x 8 block of the 64 x 64 result (4 LDS.128 per 64 FMA), the groups' blocks are added through shared memory
// at the end.  The rows per block are chosen on the host (wg_rows) so that a launch fills the GPU a few times over
// while keeping the number of partials — the fixed-order reduction that follows — small.
constexpr int WG_SUB = 64, WG_LD = 68, WG_BUF = 2 * WG_SUB * WG_LD;   // floats per stage: x rows then dz rows
constexpr int WG_SMEM = 2 * WG_BUF * (int)sizeof(float);

// packed fp32 pairs in a 64-bit register (Blackwell fma.rn.f32x2: two FMAs per issued instruction)
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

// 16-byte async copy; src_bytes = 0 zero-fills the destination
__device__ __forceinline__ void cp_async16_zfill(void *smem, const void *gmem, int src_bytes) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(sa), "l"(gmem), "r"(src_bytes));
}

__global__ void __launch_bounds__(NT)
    conv_wgrad_kernel(const float *__restrict__ x, const float *__restrict__ dz, int T, int K, int P, int bias_tap,
                      int rows, int nchunk, float *__restrict__ partial_w, float *__restrict__ partial_b) {
  extern __shared__ __align__(16) float wsm[];
  // the taps of one chunk are adjacent in launch order, so they run together and share its rows through L2
  const int j = blockIdx.x % K, b = blockIdx.y, chunk = blockIdx.x / K;
  const int t_begin = chunk * rows, t_end = min(T, t_begin + rows);
  const int g = threadIdx.x >> 6, l = threadIdx.x & 63, ci0 = (l >> 3) * 8, co0 = (l & 7) * 8;
  unsigned long long acc2[8][4];          // [input channel][output channel pair], two fp32 each
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc2[a][c] = 0ull;
  float bacc = 0.0f;   // threads < 64: bias gradient of channel threadIdx.x
  const float *xb = x + (size_t)b * T * 64, *db = dz + (size_t)b * T * 64;
  auto stage = [&](int t0, int buf) {
    float *xs = wsm + buf * WG_BUF, *ds = xs + WG_SUB * WG_LD;
    for (int i = threadIdx.x; i < WG_SUB * 16; i += NT) {
      const int r = i >> 4, c4 = (i & 15) * 4, t = t0 + r, tx = t + j - P;
      const bool okd = t < t_end, okx = okd && tx >= 0 && tx < T;
      cp_async16_zfill(&ds[r * WG_LD + c4], okd ? db + (size_t)t * 64 + c4 : db, okd ? 16 : 0);
      cp_async16_zfill(&xs[r * WG_LD + c4], okx ? xb + (size_t)tx * 64 + c4 : xb, okx ? 16 : 0);
    }
    asm volatile("cp.async.commit_group;\n");
  };
  const int nsub = (t_end - t_begin + WG_SUB - 1) / WG_SUB;
  stage(t_begin, 0);
  for (int sub = 0; sub < nsub; ++sub) {
    const int buf = sub & 1;
    if (sub + 1 < nsub) {
      stage(t_begin + (sub + 1) * WG_SUB, buf ^ 1);
      asm volatile("cp.async.wait_group 1;\n");
    } else {
      asm volatile("cp.async.wait_group 0;\n");
    }
    __syncthreads();
    const float *xs = wsm + buf * WG_BUF, *ds = xs + WG_SUB * WG_LD;
#pragma unroll 2
    for (int rr = 0; rr < WG_SUB / 4; ++rr) {
      const int r = g * (WG_SUB / 4) + rr;
      const float4 x0 = *reinterpret_cast<const float4 *>(&xs[r * WG_LD + ci0]);
      const float4 x1 = *reinterpret_cast<const float4 *>(&xs[r * WG_LD + ci0 + 4]);
      // packed fp32 FMAs (fma.rn.f32x2): the dz row is read as four 64-bit pairs, x values are duplicated
      const ulonglong2 d0 = *reinterpret_cast<const ulonglong2 *>(&ds[r * WG_LD + co0]);
      const ulonglong2 d1 = *reinterpret_cast<const ulonglong2 *>(&ds[r * WG_LD + co0 + 4]);
      const float xa[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
      const unsigned long long dp[4] = {d0.x, d0.y, d1.x, d1.y};
#pragma unroll
      for (int a = 0; a < 8; ++a) {
        const unsigned long long xx = pack2(xa[a], xa[a]);
#pragma unroll
        for (int c = 0; c < 4; ++c) acc2[a][c] = fma2x(xx, dp[c], acc2[a][c]);
      }
    }
    if (j == bias_tap && threadIdx.x < 64) {
      for (int r = 0; r < WG_SUB; ++r) bacc += ds[r * WG_LD + threadIdx.x];
    }
    __syncthreads();   // this buffer is refilled two iterations from now, by the stage() issued next iteration
  }
  // add the four row groups' blocks in a fixed order through shared memory (a 64 x 64 tile, pitch 68)
  float *tile = wsm;
  float acc[8][8];
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c) unpack2(acc2[a][c], acc[a][2 * c], acc[a][2 * c + 1]);
  for (int gg = 0; gg < 4; ++gg) {
    if (g == gg) {
#pragma unroll
      for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float *p = &tile[(ci0 + a) * WG_LD + co0 + c];
          *p = gg == 0 ? acc[a][c] : *p + acc[a][c];
        }
    }
    __syncthreads();
  }
  const size_t blk = (size_t)b * nchunk + chunk;
  float *pw = partial_w + (blk * K + j) * 4096;
  for (int i = threadIdx.x; i < 1024; i += NT) {
    const int r = i >> 4, c4 = (i & 15) * 4;
    *reinterpret_cast<float4 *>(&pw[r * 64 + c4]) = *reinterpret_cast<const float4 *>(&tile[r * WG_LD + c4]);
  }
  if (j == bias_tap && threadIdx.x < 64) partial_b[blk * 64 + threadIdx.x] = bacc;
}

// out[i] = sum_blk partial[blk * n + i], fixed order: a block owns 32 outputs, 8 interleaved segments of the partials
// are summed in parallel and then added in a fixed tree.  Launch with 256 threads, ceil(n / 32) blocks.
__global__ void __launch_bounds__(256)
    sum_partials_f_kernel(const float *__restrict__ partial, int nblk, int n, float *__restrict__ out) {
  __shared__ double red[8][32];
  const int o = threadIdx.x & 31, seg = threadIdx.x >> 5, i = blockIdx.x * 32 + o;
  double t = 0.0;
  if (i < n)
    for (int b = seg; b < nblk; b += 8) t += (double)partial[(size_t)b * n + i];
  red[seg][o] = t;
  __syncthreads();
  if (seg == 0 && i < n)
    out[i] = (float)(((red[0][o] + red[1][o]) + (red[2][o] + red[3][o])) + ((red[4][o] + red[5][o]) + (red[6][o] + red[7][o])));
}

// Conv1d(1,64,7,p=3) gradients: partial dw[blk][7][64], db[blk][64] from s[b][t], dx[b][t][64]
constexpr int IN_ROWS = 1024;
__global__ void __launch_bounds__(NT)
    conv_in_wgrad_kernel(const float *__restrict__ s, const float *__restrict__ dx, int T, int nchunk,
                         float *__restrict__ partial) {
  __shared__ float red[4][8][64];
  const int b = blockIdx.y, chunk = blockIdx.x, c = threadIdx.x & 63, part = threadIdx.x >> 6;
  const int t_begin = chunk * IN_ROWS, t_end = min(T, t_begin + IN_ROWS);
  const float *sb = s + (size_t)b * T, *db = dx + (size_t)b * T * 64;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int t = t_begin + part; t < t_end; t += 4) {
    const float d = db[(size_t)t * 64 + c];
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      const int ts = t + j - 3;
      acc[j] = fmaf((ts >= 0 && ts < T) ? sb[ts] : 0.0f, d, acc[j]);
    }
    acc[7] += d;
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[part][j][c] = acc[j];
  __syncthreads();
  if (part == 0) {
    float *p = partial + ((size_t)b * nchunk + chunk) * 512;
#pragma unroll
    for (int j = 0; j < 8; ++j) p[j * 64 + c] = red[0][j][c] + red[1][j][c] + red[2][j][c] + red[3][j][c];
  }
}

// ds[b][t] = sum_j sum_c w[j][c] dx[b][t - j + 3][c]
__global__ void conv_in_dgrad_kernel(const float *__restrict__ dx, const float *__restrict__ w, float *__restrict__ ds, int T) {
  __shared__ float ws[7 * 64];
  for (int i = threadIdx.x; i < 448; i += blockDim.x) ws[i] = w[i];
  __syncthreads();
  const int b = blockIdx.y, t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const float *db = dx + (size_t)b * T * 64;
  float a = 0.0f;
  for (int j = 0; j < 7; ++j) {
    const int tt = t - j + 3;
    if (tt < 0 || tt >= T) continue;
    const float4 *row = reinterpret_cast<const float4 *>(db + (size_t)tt * 64);
    for (int c4 = 0; c4 < 16; ++c4) {
      const float4 v = row[c4];
      a = fmaf(v.x, ws[j * 64 + c4 * 4], a); a = fmaf(v.y, ws[j * 64 + c4 * 4 + 1], a);
      a = fmaf(v.z, ws[j * 64 + c4 * 4 + 2], a); a = fmaf(v.w, ws[j * 64 + c4 * 4 + 3], a);
    }
  }
  ds[(size_t)b * T + t] = a;
}

// BCE-with-logits gradients of  lam_loc * mean(BCE(ch 0, [b < B_wm])) + lam_dec * mean over b < B_wm of BCE(ch 1+j, bit j)
__global__ void bce_heads_bwd_kernel(const float *__restrict__ logits, const long long *__restrict__ message, int B_wm,
                                     int B2, int T, int nout, float lam_loc, float lam_dec, float *__restrict__ dlog) {
  const long long rows = (long long)B2 * T;
  const float s_loc = lam_loc / (float)rows;
  const float s_dec = (B_wm > 0 && nout > 1) ? lam_dec / ((float)B_wm * (float)T * (float)(nout - 1)) : 0.0f;
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(r / T);
    const float *p = logits + r * nout;
    float *g = dlog + r * nout;
    g[0] = s_loc * (sigmoidf_(p[0]) - (b < B_wm ? 1.0f : 0.0f));
    const long long m = b < B_wm ? message[b] : 0;
    for (int j = 1; j < nout; ++j)
      g[j] = b < B_wm ? s_dec * (sigmoidf_(p[j]) - (float)((m >> (j - 1)) & 1)) : 0.0f;
  }
}

// 1x1 head (64 -> nout): dy[n][c] = sum_o dlog[n][o] w[o][c]
__global__ void __launch_bounds__(NT)
    head_dgrad_kernel(const float *__restrict__ dlog, const float *__restrict__ w, float *__restrict__ dy, long long N, int nout) {
  __shared__ float ws[WM_MAX_HEAD * 64];
  for (int i = threadIdx.x; i < nout * 64; i += NT) ws[i] = w[i];
  __syncthreads();
  const int c = threadIdx.x & 63, rl = threadIdx.x >> 6;
  for (long long r = (long long)blockIdx.x * 4 + rl; r < N; r += (long long)gridDim.x * 4) {
    float a = 0.0f;
    for (int o = 0; o < nout; ++o) a = fmaf(dlog[r * nout + o], ws[o * 64 + c], a);
    dy[r * 64 + c] = a;
  }
}

// partial dW[blk][o][c] = sum_rows dlog[r][o] y[r][c];  partial db[blk][o] = sum_rows dlog[r][o]
constexpr int HW_ROWS = 1024;
__global__ void __launch_bounds__(NT)
    head_wgrad_kernel(const float *__restrict__ dlog, const float *__restrict__ y, long long N, int nout,
                      float *__restrict__ partial_w, float *__restrict__ partial_b) {
  const int c = threadIdx.x & 63, og = threadIdx.x >> 6;   // outputs og, og + 4, ...
  const long long r0 = (long long)blockIdx.x * HW_ROWS, r1 = r0 + HW_ROWS < N ? r0 + HW_ROWS : N;
  float acc[WM_MAX_HEAD / 4], bacc[WM_MAX_HEAD / 4];
#pragma unroll
  for (int k = 0; k < WM_MAX_HEAD / 4; ++k) { acc[k] = 0.0f; bacc[k] = 0.0f; }
  for (long long r = r0; r < r1; ++r) {
    const float yv = y[r * 64 + c];
#pragma unroll
    for (int k = 0; k < WM_MAX_HEAD / 4; ++k) {
      const int o = og + 4 * k;
      if (o < nout) {
        const float d = dlog[r * nout + o];
        acc[k] = fmaf(d, yv, acc[k]);
        bacc[k] += d;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < WM_MAX_HEAD / 4; ++k) {
    const int o = og + 4 * k;
    if (o < nout) {
      partial_w[((size_t)blockIdx.x * nout + o) * 64 + c] = acc[k];
      if (c == 0) partial_b[(size_t)blockIdx.x * nout + o] = bacc[k];
    }
  }
}

// torch.optim.Adam (no weight decay, no amsgrad), bias-corrected
__global__ void adam_kernel(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m, float *__restrict__ v,
                            long long n, float lr, float b1, float b2, float eps, float bc1, float bc2) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i];
    const float mi = b1 * m[i] + (1.0f - b1) * gi;
    const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] -= lr * (mi / bc1) / (sqrtf(vi / bc2) + eps);
  }
}

// y[b][t][c] = x[b][t][c] + e[message[b]][c]                  (py/main16.py:155-157)
__global__ void add_embedding_kernel(const float *__restrict__ x, const float *__restrict__ emb,
                                     const long long *__restrict__ message, float *__restrict__ y, int T, long long n4) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const int c4 = (int)(i & 15) * 4;
    const int b = (int)(i / ((long long)T * 16));
    const float4 v = reinterpret_cast<const float4 *>(x)[i];
    const float4 e = *reinterpret_cast<const float4 *>(emb + message[b] * 64 + c4);
    reinterpret_cast<float4 *>(y)[i] = make_float4(v.x + e.x, v.y + e.y, v.z + e.z, v.w + e.w);
  }
}

// colsum[b][c] = sum_t g[b][t][c]  (one block per clip, fixed order)
__global__ void __launch_bounds__(NT) clip_colsum_kernel(const float *__restrict__ g, float *__restrict__ colsum, int T) {
  __shared__ double red[4][64];
  const int b = blockIdx.x, c = threadIdx.x & 63, part = threadIdx.x >> 6;
  const float *gb = g + (size_t)b * T * 64;
  double a = 0.0;
  for (int t = part; t < T; t += 4) a += (double)gb[(size_t)t * 64 + c];
  red[part][c] = a;
  __syncthreads();
  if (part == 0) colsum[b * 64 + c] = (float)((red[0][c] + red[1][c]) + (red[2][c] + red[3][c]));
}

// demb[message[b]][c] += colsum[b][c], clips in order (repeated messages accumulate deterministically)
__global__ void embedding_scatter_kernel(const float *__restrict__ colsum, const long long *__restrict__ message,
                                         float *__restrict__ demb, int B) {
  const int c = threadIdx.x;
  for (int b = 0; b < B; ++b) demb[message[b] * 64 + c] += colsum[b * 64 + c];
}

// losses: {l1, mel, loud, loc, bce, hf, total, raw_total}
__global__ void train_totals_kernel(float *losses, float l1, float ms, float ld, float lc, float dc, float hf) {
  losses[7] = losses[0] + losses[1] + losses[2] + losses[3] + losses[4];
  losses[6] = l1 * losses[0] + ms * losses[1] + ld * losses[2] + lc * losses[3] + dc * losses[4] + hf * losses[5];
}

int grid_for(long long n, int per_block) {
  long long g = (n + per_block - 1) / per_block;
  return (int)(g < 1 ? 1 : (g > 8LL * sm_count() ? 8LL * sm_count() : g));
}

}  // namespace

// ---- launchers (all async on st) --------------------------------------------------------------------------
// scratch: doubles [nblk * 128 + 128]; returns sums {S1, S2} at scratch + nblk * 128
static int chan_sums(int mode, const float *a, const float *b, const float *act, const float *mean, const float *rstd,
                     long long N, double *scratch, double **sums_out, cudaStream_t st) {
  const int rows_per_block = 1024;
  const int nblk = (int)((N + rows_per_block - 1) / rows_per_block);
  if (mode == 0) chan_sums_kernel<0><<<nblk, NT, 0, st>>>(a, b, act, mean, rstd, N, rows_per_block, scratch);
  else chan_sums_kernel<1><<<nblk, NT, 0, st>>>(a, b, act, mean, rstd, N, rows_per_block, scratch);
  WM_CHECK_LAUNCH("chan_sums");
  double *sums = scratch + (size_t)nblk * 128;
  sum2_kernel<<<1, 1024, 0, st>>>(scratch, nblk, sums);
  WM_CHECK_LAUNCH("sum2");
  *sums_out = sums;
  return 0;
}

size_t train_scratch_doubles(long long N) { return (size_t)((N + 1023) / 1024) * 128 + 128; }

// z -> batch statistics (mean, rstd [64] each, running stats updated) -> out = relu?(bn(z) + residual)
int launch_bn_train_fwd(const float *z, const float *gamma, const float *beta, const float *residual, float *out,
                        float *mean, float *rstd, float *run_mean, float *run_var, long long N, int relu, double *scratch,
                        cudaStream_t st) {
  double *sums = nullptr;
  WM_TRY(chan_sums(0, z, nullptr, nullptr, nullptr, nullptr, N, scratch, &sums, st));
  bn_finish_stats_kernel<<<1, 64, 0, st>>>(sums, N, 1e-5f, 0.1f, mean, rstd, run_mean, run_var);
  WM_CHECK_LAUNCH("bn_finish_stats");
  bn_apply_kernel<<<grid_for(N * 16, NT), NT, 0, st>>>(z, mean, rstd, gamma, beta, residual, out, N * 16, relu);
  WM_CHECK_LAUNCH("bn_apply");
  return 0;
}

// dout (grad of the block's post-activation output `act`, nullable = no ReLU) -> dz, dres (nullable), dgamma, dbeta
int launch_bn_train_bwd(const float *dout, const float *act, const float *z, const float *mean, const float *rstd,
                        const float *gamma, float *dz, float *dres, float *dgamma, float *dbeta, long long N,
                        double *scratch, cudaStream_t st) {
  double *sums = nullptr;
  WM_TRY(chan_sums(1, dout, z, act, mean, rstd, N, scratch, &sums, st));
  bn_bwd_apply_kernel<<<grid_for(N * 16, NT), NT, 0, st>>>(dout, act, z, mean, rstd, gamma, sums, N, dz, dres, N * 16);
  WM_CHECK_LAUNCH("bn_bwd_apply");
  bn_param_grads_kernel<<<1, 64, 0, st>>>(sums, dgamma, dbeta);
  WM_CHECK_LAUNCH("bn_param_grads");
  return 0;
}

int launch_transpose_flip(const float *w, float *wt, int K, cudaStream_t st) {
  transpose_flip_kernel<<<(K * 4096 + 255) / 256, 256, 0, st>>>(w, wt, K);
  WM_CHECK_LAUNCH("transpose_flip");
  return 0;
}

// rows of one clip per weight-gradient block: about eight blocks per SM over the launch, never below 512 rows
static int wg_rows(int B, int T, int K) {
  int per_clip = (8 * sm_count()) / (B * K > 0 ? B * K : 1);
  const int most = (T + 511) / 512;
  if (per_clip > most) per_clip = most;
  if (per_clip < 1) per_clip = 1;
  const int rows = (T + per_clip - 1) / per_clip;
  return (rows + WG_SUB - 1) / WG_SUB * WG_SUB;
}

size_t conv_wgrad_scratch_floats(int B, int T, int K) {
  const int rows = wg_rows(B, T, K);
  const size_t nblk = (size_t)B * ((T + rows - 1) / rows);
  return nblk * ((size_t)K * 4096 + 64);
}

// dW[K][64][64], db[64] of a 64->64 'same' convolution from its input x and output gradient dz
// general form: tap j pairs x[t + j - P] with dz[t]; db (nullable) sums dz on tap `bias_tap`
int launch_conv_wgrad_ex(const float *x, const float *dz, float *dw, float *db, int B, int T, int K, int P, int bias_tap,
                         float *scratch, cudaStream_t st) {
  const int rows = wg_rows(B, T, K), nchunk = (T + rows - 1) / rows, nblk = B * nchunk;
  float *pw = scratch, *pb = scratch + (size_t)nblk * K * 4096;
  static bool attr_set = false;
  if (!attr_set) {
    WM_CHECK_CUDA(cudaFuncSetAttribute(conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM));
    attr_set = true;
  }
  conv_wgrad_kernel<<<dim3(nchunk * K, B), NT, WG_SMEM, st>>>(x, dz, T, K, P, db ? bias_tap : -1, rows, nchunk, pw, pb);
  WM_CHECK_LAUNCH("conv_wgrad");
  sum_partials_f_kernel<<<(K * 4096 + 31) / 32, 256, 0, st>>>(pw, nblk, K * 4096, dw);
  WM_CHECK_LAUNCH("sum_partials(w)");
  if (db) {
    sum_partials_f_kernel<<<2, 256, 0, st>>>(pb, nblk, 64, db);
    WM_CHECK_LAUNCH("sum_partials(b)");
  }
  return 0;
}

int launch_conv_wgrad(const float *x, const float *dz, float *dw, float *db, int B, int T, int K, float *scratch,
                      cudaStream_t st) {
  return launch_conv_wgrad_ex(x, dz, dw, db, B, T, K, K / 2, K / 2, scratch, st);
}

size_t conv_in_grads_scratch_floats(int B, int T) { return (size_t)B * ((T + IN_ROWS - 1) / IN_ROWS) * 512 + 512; }
size_t head_bwd_scratch_floats(long long N, int nout) { return (size_t)((N + HW_ROWS - 1) / HW_ROWS) * (nout * 64 + nout); }

int launch_conv_in_grads(const float *s, const float *dx, const float *w, float *dw, float *db, float *ds, int B, int T,
                         float *scratch, cudaStream_t st) {
  const int nchunk = (T + IN_ROWS - 1) / IN_ROWS, nblk = B * nchunk;
  conv_in_wgrad_kernel<<<dim3(nchunk, B), NT, 0, st>>>(s, dx, T, nchunk, scratch);
  WM_CHECK_LAUNCH("conv_in_wgrad");
  sum_partials_f_kernel<<<16, 256, 0, st>>>(scratch, nblk, 512, scratch + (size_t)nblk * 512);
  WM_CHECK_LAUNCH("sum_partials(in)");
  WM_CHECK_CUDA(cudaMemcpyAsync(dw, scratch + (size_t)nblk * 512, 448 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  WM_CHECK_CUDA(cudaMemcpyAsync(db, scratch + (size_t)nblk * 512 + 448, 64 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (ds) {
    conv_in_dgrad_kernel<<<dim3((T + 255) / 256, B), 256, 0, st>>>(dx, w, ds, T);
    WM_CHECK_LAUNCH("conv_in_dgrad");
  }
  return 0;
}

int launch_bce_heads_bwd(const float *logits, const int64_t *message, int B_wm, int B2, int T, int nout, float lam_loc,
                         float lam_dec, float *dlog, cudaStream_t st) {
  bce_heads_bwd_kernel<<<grid_for((long long)B2 * T, NT), NT, 0, st>>>(logits, reinterpret_cast<const long long *>(message),
                                                                      B_wm, B2, T, nout, lam_loc, lam_dec, dlog);
  WM_CHECK_LAUNCH("bce_heads_bwd");
  return 0;
}

int launch_head_bwd(const float *dlog, const float *y, const float *w, float *dy, float *dw, float *db, long long N,
                    int nout, float *scratch, cudaStream_t st) {
  head_dgrad_kernel<<<grid_for(N, 4), NT, 0, st>>>(dlog, w, dy, N, nout);
  WM_CHECK_LAUNCH("head_dgrad");
  const int nblk = (int)((N + HW_ROWS - 1) / HW_ROWS);
  float *pw = scratch, *pb = scratch + (size_t)nblk * nout * 64;
  head_wgrad_kernel<<<nblk, NT, 0, st>>>(dlog, y, N, nout, pw, pb);
  WM_CHECK_LAUNCH("head_wgrad");
  sum_partials_f_kernel<<<(nout * 64 + 31) / 32, 256, 0, st>>>(pw, nblk, nout * 64, dw);
  WM_CHECK_LAUNCH("sum_partials(hw)");
  sum_partials_f_kernel<<<(nout + 31) / 32, 256, 0, st>>>(pb, nblk, nout, db);
  WM_CHECK_LAUNCH("sum_partials(hb)");
  return 0;
}

int launch_adam(float *p, const float *g, float *m, float *v, long long n, float lr, float b1, float b2, float eps,
                int step, cudaStream_t st) {
  const float bc1 = 1.0f - powf(b1, (float)step), bc2 = 1.0f - powf(b2, (float)step);
  adam_kernel<<<grid_for(n, NT), NT, 0, st>>>(p, g, m, v, n, lr, b1, b2, eps, bc1, bc2);
  WM_CHECK_LAUNCH("adam");
  return 0;
}

// ---- 64 -> 64 convolutions of the training step (forward, data gradient, weight gradient) ---------------------
// In the library's default math mode (WM_MATH_BF16X2) the channel-heavy convolutions of the training step run on the
// tensor cores, like the inference path: bf16 hi+lo operand pairs, fp32 accumulation in TMEM.
//   forward / data gradient: the inference conv kernel (three partial products).  The fp32 channels-last input is
//     re-laid as hi/lo planes (one HBM pass), the weights -- which change every step -- are re-imaged by a
//     12 K-element kernel, the output comes back fp32 channels-last with the fp32 residual added in the epilogue;
//   weight gradient: wm_wgrad_tc.cu (MN-major operands straight from the planes, four partial products).
// Precision: 2^-17 relative per operand, 64x finer than the TF32 the reference itself trains with
// (torch.set_float32_matmul_precision('high'), py/main16.py:44, and cuDNN's TF32 default); measured against the fp64
// oracle next to PyTorch's own step in profiles/r2_train_precision_tc.txt.  WM_MATH_FP32 (wm_set_math_mode) selects the
// exact-order fp32 FMA kernels, against which the tight operator-level gradient tests run.
namespace {
struct TcScratch { void *xp, *yp; float *img, *wg; };
size_t tc_planar_floats(int B, int T) { return ((size_t)B * ((size_t)T + 2 * WM_PLANAR_PAD) * 64 + 1024 + 63) / 64 * 64; }
constexpr size_t kTcImgFloats = 7 * 4096;
size_t tc_scratch_floats(int B, int T) {
  return 2 * tc_planar_floats(B, T) + kTcImgFloats + wgrad_tc_scratch_floats(7) + colsum64_scratch_floats() + 256;
}
bool train_tc_enabled() { return math_mode() == WM_MATH_BF16X2; }
// y = conv(in) + bias + residual   (w tap-major [K][ci][co])
int conv64_train(const float *in, const float *w, const float *bias, const float *residual, float *out, int B, int T, int K,
                 TcScratch tc, cudaStream_t st) {
  if (!train_tc_enabled() || tc.xp == nullptr || (K != 3 && K != 7))
    return launch_conv64_fp32(in, w, bias, residual, nullptr, out, B, T, K, 0, st);
  WM_TRY(launch_to_planar(in, nullptr, tc.xp, B, T, st));
  WM_TRY(launch_pack_conv64_tc(w, tc.img, K, st));
  return launch_conv64_tc(tc.xp, tc.img, bias, nullptr, nullptr, out, B, T, K, 0, st, residual);
}
// gradients of y = conv(x): dw, db, and dx = conv^T(dy) + dx_residual (dx nullable); wt [K*4096] and zero64 scratch
int conv64_bwd_train(const float *x, const float *dy, const float *w, float *dw, float *db, float *dx,
                     const float *dx_residual, int B, int T, int K, TcScratch tc, float *fscratch, float *wt,
                     const float *zero64, cudaStream_t st) {
  if (!train_tc_enabled() || tc.xp == nullptr || (K != 3 && K != 7)) {
    WM_TRY(launch_conv_wgrad(x, dy, dw, db, B, T, K, fscratch, st));
    if (dx == nullptr) return 0;
    WM_TRY(launch_transpose_flip(w, wt, K, st));
    return launch_conv64_fp32(dy, wt, zero64, dx_residual, nullptr, dx, B, T, K, 0, st);
  }
  WM_TRY(launch_to_planar(dy, nullptr, tc.yp, B, T, st));
  WM_TRY(launch_to_planar(x, nullptr, tc.xp, B, T, st));
  WM_TRY(launch_wgrad_tc(tc.xp, tc.yp, dw, B, T, K, tc.wg, st));
  WM_TRY(launch_colsum64(dy, db, (long long)B * T, tc.wg + wgrad_tc_scratch_floats(7), st));
  if (dx == nullptr) return 0;
  WM_TRY(launch_transpose_flip(w, wt, K, st));
  WM_TRY(launch_pack_conv64_tc(wt, tc.img, K, st));
  return launch_conv64_tc(tc.yp, tc.img, zero64, nullptr, nullptr, dx, B, T, K, 0, st, dx_residual);
}
TcScratch tc_take(float *base, size_t &off, int B, int T) {   // carve the scratch out of a float workspace (base may be null)
  TcScratch t;
  auto take = [&](size_t n) { float *q = base ? base + off : nullptr; off += (n + 63) / 64 * 64; return q; };
  t.xp = take(tc_planar_floats(B, T));
  t.yp = take(tc_planar_floats(B, T));
  t.img = take(kTcImgFloats);
  t.wg = take(wgrad_tc_scratch_floats(7) + colsum64_scratch_floats() + 64);
  return t;
}
}  // namespace

// single-operator forms behind the C ABI (wm_conv64_train_fwd, wm_conv64_bwd): same code path as inside the step
size_t train_conv64_scratch_floats(int B, int T, int K) {
  return tc_scratch_floats(B, T) + conv_wgrad_scratch_floats(B, T, K) + (size_t)K * 4096 + 128;
}
int train_conv64_fwd(const float *x, const float *w, const float *bias, const float *residual, float *y, int B, int T,
                     int K, float *scratch, cudaStream_t st) {
  size_t off = 0;
  TcScratch tc = tc_take(scratch, off, B, T);
  return conv64_train(x, w, bias, residual, y, B, T, K, tc, st);
}
int train_conv64_bwd(const float *x, const float *dy, const float *w, float *dw, float *db, float *dx, int B, int T, int K,
                     float *scratch, cudaStream_t st) {
  size_t off = 0;
  TcScratch tc = tc_take(scratch, off, B, T);
  float *fscratch = scratch + off;
  float *wt = fscratch + conv_wgrad_scratch_floats(B, T, K), *zero = wt + (size_t)K * 4096;
  WM_CHECK_CUDA(cudaMemsetAsync(zero, 0, 64 * sizeof(float), st));
  return conv64_bwd_train(x, dy, w, dw, db, dx, nullptr, B, T, K, tc, fscratch, wt, zero, st);
}

// ---- the detector's training step ---------------------------------------------------------------------------
namespace {
struct DetWs {
  float *x0, *act[2][4] /* z1, u, z2, y */, *logits, *dlog, *g[3], *wt, *stats, *zero64, *fscratch, *losses;
  double *dscratch;
  TcScratch tc;
  size_t bytes;
};
size_t align64(size_t n) { return (n + 63) / 64 * 64; }
DetWs det_ws(void *base, int B2, int T, int nout) {
  const size_t N = (size_t)B2 * T, A = align64(N * 64), L = align64(N * nout);
  const size_t nb_h = (N + HW_ROWS - 1) / HW_ROWS, nb_in = (size_t)B2 * ((T + IN_ROWS - 1) / IN_ROWS);
  size_t fs = conv_wgrad_scratch_floats(B2, T, 3);
  if (fs < nb_h * (nout * 64 + nout)) fs = nb_h * (nout * 64 + nout);
  if (fs < nb_in * 512 + 512) fs = nb_in * 512 + 512;
  const size_t loss_f = wm_loss_workspace_bytes(B2, T) / sizeof(float);
  if (fs < loss_f) fs = loss_f;
  DetWs w;
  float *p = (float *)base;
  size_t off = 0;
  auto take = [&](size_t n) { float *q = p ? p + off : nullptr; off += align64(n); return q; };
  w.x0 = take(A);
  for (int k = 0; k < 2; ++k)
    for (int i = 0; i < 4; ++i) w.act[k][i] = take(A);
  w.logits = take(L);
  w.dlog = take(L);
  for (int i = 0; i < 3; ++i) w.g[i] = take(A);
  w.wt = take(3 * 4096);
  w.stats = take(4 * 128);
  w.zero64 = take(64);
  w.losses = take(64);
  w.fscratch = take(fs);
  w.dscratch = (double *)take(2 * train_scratch_doubles((long long)N));
  w.tc = tc_take(p, off, B2, T);
  w.bytes = off * sizeof(float);
  return w;
}
}  // namespace

size_t detector_train_workspace_bytes(int B2, int T, int nout) { return det_ws(nullptr, B2, T, nout).bytes; }

// One training step of the Detector (py/main16.py:160-176 in train mode, losses of :249-264) on x[B2][T] whose first
// B_wm clips carry message[b]:  loss = lam_loc * loc + lam_dec * bce.  params / grads / adam_m / adam_v: WM_DT_SIZE
// floats; run_stats: WM_DT_STATS floats (running mean, var of the four BatchNorms); losses_out: {loc, bce} (device).
// d_input (nullable): gradient of the loss w.r.t. x.  adam_step = 0 skips the optimizer (gradients only).
int detector_train_step(float *params, float *grads, float *adam_m, float *adam_v, float *run_stats, const float *x,
                        const int64_t *message, int B_wm, int B2, int T, int nout, float lam_loc, float lam_dec, float lr,
                        float beta1, float beta2, float eps, int adam_step, float *losses_out, float *d_input,
                        void *workspace, cudaStream_t st) {
  DetWs w = det_ws(workspace, B2, T, nout);
  const long long N = (long long)B2 * T;
  WM_CHECK_CUDA(cudaMemsetAsync(w.zero64, 0, 64 * sizeof(float), st));
  WM_CHECK_CUDA(cudaMemsetAsync(w.losses, 0, 2 * sizeof(float), st));
  WM_CHECK_CUDA(cudaMemsetAsync(grads, 0, (size_t)WM_DT_SIZE * sizeof(float), st));
  // forward
  WM_TRY(launch_conv_in_k7(x, params + WM_DT_IN_W, params + WM_DT_IN_B, w.x0, B2, T, st));
  const float *in = w.x0;
  for (int k = 0; k < 2; ++k) {
    const float *rb = params + WM_DT_RB0 + k * WM_DT_RB_SIZE;
    float *rs = run_stats + k * 256, *stt = w.stats + k * 256;
    float *z1 = w.act[k][0], *u = w.act[k][1], *z2 = w.act[k][2], *y = w.act[k][3];
    WM_TRY(conv64_train(in, rb + WM_DT_RB_W1, rb + WM_DT_RB_B1, nullptr, z1, B2, T, 3, w.tc, st));
    WM_TRY(launch_bn_train_fwd(z1, rb + WM_DT_RB_G1, rb + WM_DT_RB_BE1, nullptr, u, stt, stt + 64, rs, rs + 64, N, 1,
                               w.dscratch, st));
    WM_TRY(conv64_train(u, rb + WM_DT_RB_W2, rb + WM_DT_RB_B2, nullptr, z2, B2, T, 3, w.tc, st));
    WM_TRY(launch_bn_train_fwd(z2, rb + WM_DT_RB_G2, rb + WM_DT_RB_BE2, in, y, stt + 128, stt + 192, rs + 128, rs + 192, N,
                               1, w.dscratch, st));
    in = y;
  }
  WM_TRY(launch_head(in, params + WM_DT_HEAD_W, params + WM_DT_HEAD_B, w.logits, B2, T, nout, st));
  WM_TRY(launch_bce_heads(w.logits, message, w.losses, nout > 1 ? w.losses + 1 : nullptr, w.fscratch, B_wm, B2, T, nout, st));
  if (losses_out) WM_CHECK_CUDA(cudaMemcpyAsync(losses_out, w.losses, 2 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  // backward
  WM_TRY(launch_bce_heads_bwd(w.logits, message, B_wm, B2, T, nout, lam_loc, lam_dec, w.dlog, st));
  float *gA = w.g[0], *gB = w.g[1], *gC = w.g[2];
  WM_TRY(launch_head_bwd(w.dlog, in, params + WM_DT_HEAD_W, gA, grads + WM_DT_HEAD_W, grads + WM_DT_HEAD_B, N, nout,
                         w.fscratch, st));
  for (int k = 1; k >= 0; --k) {
    const float *rb = params + WM_DT_RB0 + k * WM_DT_RB_SIZE;
    float *gr = grads + WM_DT_RB0 + k * WM_DT_RB_SIZE;
    const float *stt = w.stats + k * 256;
    const float *xin = k == 0 ? w.x0 : w.act[0][3];
    const float *z1 = w.act[k][0], *u = w.act[k][1], *z2 = w.act[k][2], *y = w.act[k][3];
    WM_TRY(launch_bn_train_bwd(gA, y, z2, stt + 128, stt + 192, rb + WM_DT_RB_G2, gB, gC, gr + WM_DT_RB_G2,
                               gr + WM_DT_RB_BE2, N, w.dscratch, st));
    WM_TRY(conv64_bwd_train(u, gB, rb + WM_DT_RB_W2, gr + WM_DT_RB_W2, gr + WM_DT_RB_B2, gA, nullptr, B2, T, 3, w.tc,
                            w.fscratch, w.wt, w.zero64, st));
    WM_TRY(launch_bn_train_bwd(gA, u, z1, stt, stt + 64, rb + WM_DT_RB_G1, gB, nullptr, gr + WM_DT_RB_G1, gr + WM_DT_RB_BE1,
                               N, w.dscratch, st));
    WM_TRY(conv64_bwd_train(xin, gB, rb + WM_DT_RB_W1, gr + WM_DT_RB_W1, gr + WM_DT_RB_B1, gA, gC, B2, T, 3, w.tc,
                            w.fscratch, w.wt, w.zero64, st));
  }
  WM_TRY(launch_conv_in_grads(x, gA, params + WM_DT_IN_W, grads + WM_DT_IN_W, grads + WM_DT_IN_B, d_input, B2, T,
                              w.fscratch, st));
  if (adam_step > 0)
    WM_TRY(launch_adam(params, grads, adam_m, adam_v, WM_DT_SIZE, lr, beta1, beta2, eps, adam_step, st));
  return 0;
}

// ---- the whole training step: forward and backward of py/main16.py:244-277 ------------------------------------
namespace {

struct RbActs { float *z1, *u, *z2, *y; };

// relu(x + BN(conv3(relu(BN(conv3(x)))))) in train mode; rb: WM_DT_RB_* block, rs: 256 running stats, stt: 256 batch stats
int rb_train_fwd(const float *in, const float *rb, float *rs, float *stt, RbActs a, int B, int T, double *dscratch,
                 TcScratch tc, cudaStream_t st) {
  const long long N = (long long)B * T;
  WM_TRY(conv64_train(in, rb + WM_DT_RB_W1, rb + WM_DT_RB_B1, nullptr, a.z1, B, T, 3, tc, st));
  WM_TRY(launch_bn_train_fwd(a.z1, rb + WM_DT_RB_G1, rb + WM_DT_RB_BE1, nullptr, a.u, stt, stt + 64, rs, rs + 64, N, 1,
                             dscratch, st));
  WM_TRY(conv64_train(a.u, rb + WM_DT_RB_W2, rb + WM_DT_RB_B2, nullptr, a.z2, B, T, 3, tc, st));
  return launch_bn_train_fwd(a.z2, rb + WM_DT_RB_G2, rb + WM_DT_RB_BE2, in, a.y, stt + 128, stt + 192, rs + 128, rs + 192,
                             N, 1, dscratch, st);
}

// gA holds dL/dy on entry and dL/dx on exit; gB, gC scratch activations; wt [3*4096], zero64 as in DetWs
int rb_train_bwd(const float *xin, const float *rb, float *gr, const float *stt, RbActs a, float *gA, float *gB, float *gC,
                 float *wt, const float *zero64, int B, int T, float *fscratch, double *dscratch, TcScratch tc,
                 cudaStream_t st) {
  const long long N = (long long)B * T;
  WM_TRY(launch_bn_train_bwd(gA, a.y, a.z2, stt + 128, stt + 192, rb + WM_DT_RB_G2, gB, gC, gr + WM_DT_RB_G2,
                             gr + WM_DT_RB_BE2, N, dscratch, st));
  WM_TRY(conv64_bwd_train(a.u, gB, rb + WM_DT_RB_W2, gr + WM_DT_RB_W2, gr + WM_DT_RB_B2, gA, nullptr, B, T, 3, tc, fscratch,
                          wt, zero64, st));
  WM_TRY(launch_bn_train_bwd(gA, a.u, a.z1, stt, stt + 64, rb + WM_DT_RB_G1, gB, nullptr, gr + WM_DT_RB_G1,
                             gr + WM_DT_RB_BE1, N, dscratch, st));
  return conv64_bwd_train(xin, gB, rb + WM_DT_RB_W1, gr + WM_DT_RB_W1, gr + WM_DT_RB_B1, gA, gC, B, T, 3, tc, fscratch, wt,
                          zero64, st);
}

struct GenWs {
  float *x0, *h, *hE, *ct, *gates, *cell, *g[3], *wt, *stats, *zero64, *losses, *colsum;
  RbActs rb[3];
  float *draw, *d1, *delta, *xdet /* [2B][T]: s_w then s */, *gsw, *gdraw, *dxdet, *fscratch, *lstm_scratch;
  double *dscratch;
  TcScratch tc;
  void *det_ws;
  size_t bytes;
};

GenWs gen_ws(void *base, int B, int T, int nout) {
  const size_t N = (size_t)B * T, A = align64(N * 64), V = align64(N);
  size_t fs = conv_wgrad_scratch_floats(B, T, 7);
  const size_t nb_h = (N + HW_ROWS - 1) / HW_ROWS, nb_in = (size_t)B * ((T + IN_ROWS - 1) / IN_ROWS);
  if (fs < nb_h * 65) fs = nb_h * 65;
  if (fs < nb_in * 512 + 512) fs = nb_in * 512 + 512;
  const size_t loss_f = wm_loss_workspace_bytes(B, T) / sizeof(float);
  if (fs < loss_f) fs = loss_f;
  size_t gf = stft_bwd_scratch_floats(B, T, 2048, 512);
  if (gf < stft_bwd_scratch_floats(B, T, 1024, 256)) gf = stft_bwd_scratch_floats(B, T, 1024, 256);
  if (gf < stft_bwd_scratch_floats(B, T, 512, 128)) gf = stft_bwd_scratch_floats(B, T, 512, 128);
  if (fs < gf) fs = gf;
  if (fs < N) fs = N;
  GenWs w;
  float *p = (float *)base;
  size_t off = 0;
  auto take = [&](size_t n) { float *q = p ? p + off : nullptr; off += align64(n); return q; };
  w.x0 = take(A);
  for (int k = 0; k < 3; ++k) { w.rb[k].z1 = take(A); w.rb[k].u = take(A); w.rb[k].z2 = take(A); w.rb[k].y = take(A); }
  w.h = take(A); w.hE = take(A); w.ct = take(A);
  w.gates = take(4 * A); w.cell = take(A);
  for (int i = 0; i < 3; ++i) w.g[i] = take(A);
  w.wt = take(7 * 4096);
  w.stats = take(3 * 256);
  w.zero64 = take(64);
  w.losses = take(64);
  w.colsum = take((size_t)B * 64);
  w.draw = take(V); w.d1 = take(V); w.delta = take(V); w.xdet = take(2 * V); w.gsw = take(V); w.gdraw = take(V);
  w.dxdet = take(2 * V);
  w.fscratch = take(fs);
  w.lstm_scratch = take(lstm_train_bwd_scratch_floats(B, T));
  w.dscratch = (double *)take(2 * train_scratch_doubles((long long)N));
  w.tc = tc_take(p, off, B, T);
  w.det_ws = p ? (void *)(p + off) : nullptr;
  off += align64(detector_train_workspace_bytes(2 * B, T, nout) / sizeof(float) + 64);
  w.bytes = off * sizeof(float);
  return w;
}

}  // namespace

size_t train_step_workspace_bytes(int B, int T, int nout) { return gen_ws(nullptr, B, T, nout).bytes; }

// Forward + backward of one train_one_epoch iteration (py/main16.py:244-277) on s[B][T], message[B]; gradients of
// every parameter land in g_grads (WM_GT_SIZE, embedding rows included, dense) and d_grads (WM_DT_SIZE); the caller
// all-reduces them if it wants to and applies Adam (wm_adam_step) to both.  losses_out[8] (device):
// {l1, mel, loud, loc, bce, hf, total, raw_total}.  lam[6] = {l1, msspec, loud, loc, dec, hf}.
int train_forward_backward(const float *g_params, float *g_grads, float *g_stats, const float *d_params, float *d_grads,
                           float *d_stats, const float *s, const int64_t *message, const float *fir, const float *mel_fb,
                           const int *mel_band, int n_mels, const float *lam, int B, int T, int nout, float *losses_out,
                           float *s_w_out, void *workspace, cudaStream_t st) {
  GenWs w = gen_ws(workspace, B, T, nout);
  const long long N = (long long)B * T;
  const size_t V = (size_t)N;
  const long long *msg = reinterpret_cast<const long long *>(message);
  WM_CHECK_CUDA(cudaMemsetAsync(w.zero64, 0, 64 * sizeof(float), st));
  WM_CHECK_CUDA(cudaMemsetAsync(w.losses, 0, 8 * sizeof(float), st));
  WM_CHECK_CUDA(cudaMemsetAsync(g_grads, 0, (size_t)WM_GT_SIZE * sizeof(float), st));
  // ---- generator forward (py/main16.py:148-162 in train mode) ----
  WM_TRY(launch_conv_in_k7(s, g_params + WM_GT_IN_W, g_params + WM_GT_IN_B, w.x0, B, T, st));
  WM_TRY(rb_train_fwd(w.x0, g_params + WM_GT_RB0, g_stats, w.stats, w.rb[0], B, T, w.dscratch, w.tc, st));
  WM_TRY(rb_train_fwd(w.rb[0].y, g_params + WM_GT_RB1, g_stats + 256, w.stats + 256, w.rb[1], B, T, w.dscratch, w.tc, st));
  WM_TRY(launch_lstm_train_fwd(w.rb[1].y, g_params + WM_GT_LSTM_WIH, g_params + WM_GT_LSTM_WHH, g_params + WM_GT_LSTM_BIH,
                               g_params + WM_GT_LSTM_BHH, w.h, w.gates, w.cell, B, T, st));
  add_embedding_kernel<<<grid_for(N * 16, NT), NT, 0, st>>>(w.h, g_params + WM_GT_EMB, msg, w.hE, T, N * 16);
  WM_CHECK_LAUNCH("add_embedding");
  WM_TRY(conv64_train(w.hE, g_params + WM_GT_CT_W, g_params + WM_GT_CT_B, nullptr, w.ct, B, T, 7, w.tc, st));
  WM_TRY(rb_train_fwd(w.ct, g_params + WM_GT_RB2, g_stats + 512, w.stats + 512, w.rb[2], B, T, w.dscratch, w.tc, st));
  WM_TRY(launch_head(w.rb[2].y, g_params + WM_GT_HEAD_W, g_params + WM_GT_HEAD_B, w.draw, B, T, 1, st));
  // ---- post-processing (py/main16.py:245-248): d1 = fir(delta_raw) kept for the backward ----
  WM_TRY(launch_postprocess(w.draw, nullptr, fir, w.d1, nullptr, nullptr, B, T, WM_POST_FIR, 0.02f, 0.005f, 1e-8f, st));
  WM_TRY(launch_postprocess(w.draw, s, fir, w.delta, w.xdet, nullptr, B, T, WM_POST_ALL, 0.02f, 0.005f, 1e-8f, st));
  WM_CHECK_CUDA(cudaMemcpyAsync(w.xdet + V, s, V * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (s_w_out) WM_CHECK_CUDA(cudaMemcpyAsync(s_w_out, w.xdet, V * sizeof(float), cudaMemcpyDeviceToDevice, st));
  // ---- detector on cat(s_w, s): losses, its gradients, and dL/d s_w (py/main16.py:249-264) ----
  WM_TRY(detector_train_step(const_cast<float *>(d_params), d_grads, nullptr, nullptr, d_stats, w.xdet, message, B, 2 * B, T,
                             nout, lam[3], lam[4], 0.f, 0.f, 0.f, 0.f, 0, w.losses + 3, w.dxdet, w.det_ws, st));
  // ---- the perceptual losses and their gradients (py/main16.py:266-276) ----
  const float *sw = w.xdet;
  WM_TRY(launch_abs_mean(w.delta, N, w.losses + 0, w.fscratch, st));
  WM_TRY(launch_mel_log_l1(s, sw, mel_fb, mel_band, n_mels, w.losses + 1, w.fscratch, B, T, 1024, 256, st));
  WM_TRY(launch_loudness(s, sw, w.losses + 2, w.fscratch, B, T, 2048, 512, 0.01f, st));
  WM_TRY(launch_hf_penalty(w.delta, w.losses + 5, w.fscratch, B, T, 512, 113, st));
  WM_CHECK_CUDA(cudaMemcpyAsync(w.gsw, w.dxdet, V * sizeof(float), cudaMemcpyDeviceToDevice, st));
  WM_TRY(launch_mel_log_l1_bwd(s, sw, mel_fb, mel_band, n_mels, w.gsw, w.fscratch, B, T, 1024, 256, lam[1], 1, st));
  WM_TRY(launch_loudness_bwd(s, sw, w.gsw, w.fscratch, B, T, 2048, 512, 0.01f, lam[2], 1, st));
  WM_TRY(launch_abs_mean_bwd(w.delta, w.gsw, N, lam[0], 1, st));            // s_w = s + delta: same gradient buffer
  WM_TRY(launch_hf_penalty_bwd(w.delta, w.gsw, w.fscratch, B, T, 512, 113, lam[5], 1, st));
  WM_TRY(launch_postprocess_bwd(w.gsw, w.d1, fir, w.gdraw, w.fscratch, B, T, WM_POST_ALL, 0.02f, 0.005f, 1e-8f, st));
  // ---- generator backward ----
  float *gA = w.g[0], *gB = w.g[1], *gC = w.g[2];
  WM_TRY(launch_head_bwd(w.gdraw, w.rb[2].y, g_params + WM_GT_HEAD_W, gA, g_grads + WM_GT_HEAD_W, g_grads + WM_GT_HEAD_B, N,
                         1, w.fscratch, st));
  WM_TRY(rb_train_bwd(w.ct, g_params + WM_GT_RB2, g_grads + WM_GT_RB2, w.stats + 512, w.rb[2], gA, gB, gC, w.wt, w.zero64,
                      B, T, w.fscratch, w.dscratch, w.tc, st));
  WM_TRY(conv64_bwd_train(w.hE, gA, g_params + WM_GT_CT_W, g_grads + WM_GT_CT_W, g_grads + WM_GT_CT_B, gB, nullptr, B, T, 7,
                          w.tc, w.fscratch, w.wt, w.zero64, st));     // gB = dL/d(h + e)
  clip_colsum_kernel<<<B, NT, 0, st>>>(gB, w.colsum, T);
  WM_CHECK_LAUNCH("clip_colsum");
  embedding_scatter_kernel<<<1, 64, 0, st>>>(w.colsum, msg, g_grads + WM_GT_EMB, B);
  WM_CHECK_LAUNCH("embedding_scatter");
  WM_TRY(launch_lstm_train_bwd(gB, w.rb[1].y, w.h, g_params + WM_GT_LSTM_WIH, g_params + WM_GT_LSTM_WHH, w.gates, w.cell, gA,
                               g_grads + WM_GT_LSTM_WIH, g_grads + WM_GT_LSTM_WHH, g_grads + WM_GT_LSTM_BIH, B, T,
                               w.lstm_scratch, st));
  WM_CHECK_CUDA(cudaMemcpyAsync(g_grads + WM_GT_LSTM_BHH, g_grads + WM_GT_LSTM_BIH, 256 * sizeof(float),
                                cudaMemcpyDeviceToDevice, st));
  WM_TRY(rb_train_bwd(w.rb[0].y, g_params + WM_GT_RB1, g_grads + WM_GT_RB1, w.stats + 256, w.rb[1], gA, gB, gC, w.wt,
                      w.zero64, B, T, w.fscratch, w.dscratch, w.tc, st));
  WM_TRY(rb_train_bwd(w.x0, g_params + WM_GT_RB0, g_grads + WM_GT_RB0, w.stats, w.rb[0], gA, gB, gC, w.wt, w.zero64, B, T,
                      w.fscratch, w.dscratch, w.tc, st));
  WM_TRY(launch_conv_in_grads(s, gA, g_params + WM_GT_IN_W, g_grads + WM_GT_IN_W, g_grads + WM_GT_IN_B, nullptr, B, T,
                              w.fscratch, st));
  // totals (py/main16.py:273-276)
  train_totals_kernel<<<1, 1, 0, st>>>(w.losses, lam[0], lam[1], lam[2], lam[3], lam[4], lam[5]);
  WM_CHECK_LAUNCH("train_totals");
  if (losses_out) WM_CHECK_CUDA(cudaMemcpyAsync(losses_out, w.losses, 8 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return 0;
}

}  // namespace wm
