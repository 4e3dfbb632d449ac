// CUDA-core kernels for the thin layers of the main14b_2 stack (py/main14b_2.py:107-182) that have too few channels
// for a tensor-core GEMM: the Generator's last ResidualBlock(8, 8) + final_conv_dec Conv1d(8, 1, 7) + crop as ONE
// kernel reading the planar output of the last transposed convolution, and the 2-layer LSTM(32, 32) with its weights
// in registers.  Both are bound by memory / latency, not arithmetic (15 GFLOP and 1.7 GFLOP per 1024 clips).
#include "wm_common.h"
#include "wm_tc.cuh"

namespace wm {

using namespace tc;

namespace {

constexpr int GAP = WM_PC_GAP;
constexpr int TT = 240;   // outputs per block of the tail kernel: with their halos u (TT + 8) and z (TT + 6) fit 64 x 4
constexpr int TN = 64;    // threads: four consecutive time steps each

// exp(v) - 1 through ex2.approx: absolute error ~1e-7 (expm1f is ~40 instructions, 16 of them per output sample here)
__device__ __forceinline__ float elu_f(float v) { return v > 0.0f ? v : ex2_approx(v * 1.4426950408889634f) - 1.0f; }

// a[pos][co pair] += sum over ci, k of in[ci][4 j + pos + k] * w[k][ci][co] for four consecutive positions: every weight
// vector and input value is fetched from shared memory once per FOUR outputs (one position per thread made the kernel
// LSU-bound: 72 shared loads per 192 FMAs), and the FMAs are packed f32x2 (two output channels per instruction).
template <int W>
__device__ __forceinline__ void conv8k3x4(const float (*in)[W], const float (*w)[8][8], int j, f32x2 (&a)[4][4]) {
#pragma unroll 2
  for (int ci = 0; ci < 8; ++ci) {
    const float4 v = *reinterpret_cast<const float4 *>(&in[ci][4 * j]);
    const float2 v2 = *reinterpret_cast<const float2 *>(&in[ci][4 * j + 4]);
    const f32x2 xin[6] = {pk2(v.x, v.x), pk2(v.y, v.y), pk2(v.z, v.z), pk2(v.w, v.w), pk2(v2.x, v2.x), pk2(v2.y, v2.y)};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const ulonglong2 wa = *reinterpret_cast<const ulonglong2 *>(&w[k][ci][0]);
      const ulonglong2 wb = *reinterpret_cast<const ulonglong2 *>(&w[k][ci][4]);
#pragma unroll
      for (int pos = 0; pos < 4; ++pos) {
        a[pos][0] = fma2(xin[pos + k], wa.x, a[pos][0]);
        a[pos][1] = fma2(xin[pos + k], wa.y, a[pos][1]);
        a[pos][2] = fma2(xin[pos + k], wb.x, a[pos][2]);
        a[pos][3] = fma2(xin[pos + k], wb.y, a[pos][3]);
      }
    }
  }
}

// x planar (8 channels: plane 0 = hi, plane 1 = lo; geometry (B, Tx)) ->
//   u = elu(conv1(x)), z = elu(conv2(u) + x)   (ResidualBlock, py/main14b_2.py:97-105; k3, padding 1)
//   delta[b][t] = final(z)[t] for t < T         (Conv1d(8, 1, 7, padding 3), :149,173; crop :175-177)
// shared arrays: xs[i] = x at t0 - 5 + i, us[i] = u at t0 - 4 + i, zs[i] = z at t0 - 3 + i
__global__ void __launch_bounds__(TN)
    m14_tail8_kernel(const uint4 *__restrict__ x, long long plane_rows, int Tx, const float *__restrict__ w1,
                     const float *__restrict__ b1, const float *__restrict__ w2, const float *__restrict__ b2,
                     const float *__restrict__ wf, const float *__restrict__ bf, float *__restrict__ delta, int T) {
  constexpr int WX = 4 * TN + 8, WU = 4 * TN + 8, WZ = 4 * TN + 8;   // 264: every thread reads 4 j .. 4 j + 9 at most
  __shared__ __align__(16) float xs[8][WX], us[8][WU], zs[8][WZ];
  __shared__ __align__(16) float w1s[3][8][8], w2s[3][8][8], wfs[7][8], bs[2][8];   // [tap][ci][co]
  const int b = blockIdx.y, t0 = blockIdx.x * TT, tid = threadIdx.x;
  for (int i = tid; i < 192; i += TN) {
    const int co = i & 7, ci = (i >> 3) & 7, k = i >> 6;
    w1s[k][ci][co] = w1[(co * 8 + ci) * 3 + k];
    w2s[k][ci][co] = w2[(co * 8 + ci) * 3 + k];
  }
  if (tid < 56) wfs[tid >> 3][tid & 7] = wf[(tid & 7) * 7 + (tid >> 3)];
  if (tid < 8) { bs[0][tid] = b1[tid]; bs[1][tid] = b2[tid]; }
  const long long row0 = (long long)b * (Tx + GAP) + GAP;
  for (int i = tid; i < WX; i += TN) {
    const int t = t0 - 5 + i;
    float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (t >= 0 && t < Tx) {
      const uint4 hi = __ldg(x + row0 + t), lo = __ldg(x + plane_rows + row0 + t);
      join8(hi, lo, v);
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) xs[c][i] = v[c];
  }
  __syncthreads();
  {                                               // u at indices 4 tid .. 4 tid + 3 (TT + 8 needed)
    const int j = tid;
    f32x2 a2[4][4];
    const ulonglong2 ba = *reinterpret_cast<const ulonglong2 *>(&bs[0][0]), bb = *reinterpret_cast<const ulonglong2 *>(&bs[0][4]);
#pragma unroll
    for (int pos = 0; pos < 4; ++pos) { a2[pos][0] = ba.x; a2[pos][1] = ba.y; a2[pos][2] = bb.x; a2[pos][3] = bb.y; }
    conv8k3x4<WX>(xs, w1s, j, a2);
#pragma unroll
    for (int c2 = 0; c2 < 4; ++c2) {
      float4 o0, o1;
      float *p0 = &o0.x, *p1 = &o1.x;
#pragma unroll
      for (int pos = 0; pos < 4; ++pos) {
        const int t = t0 - 4 + 4 * j + pos;
        const bool in = t >= 0 && t < Tx;
        float e0, e1;
        upk2(a2[pos][c2], e0, e1);
        p0[pos] = in ? elu_f(e0) : 0.0f;
        p1[pos] = in ? elu_f(e1) : 0.0f;
      }
      *reinterpret_cast<float4 *>(&us[2 * c2][4 * j]) = o0;
      *reinterpret_cast<float4 *>(&us[2 * c2 + 1][4 * j]) = o1;
    }
  }
  __syncthreads();
  {                                               // z at indices 4 tid .. 4 tid + 3 (TT + 6 needed)
    const int j = tid;
    f32x2 a2[4][4];
    const ulonglong2 ba = *reinterpret_cast<const ulonglong2 *>(&bs[1][0]), bb = *reinterpret_cast<const ulonglong2 *>(&bs[1][4]);
#pragma unroll
    for (int pos = 0; pos < 4; ++pos) { a2[pos][0] = ba.x; a2[pos][1] = ba.y; a2[pos][2] = bb.x; a2[pos][3] = bb.y; }
    conv8k3x4<WU>(us, w2s, j, a2);
#pragma unroll
    for (int c2 = 0; c2 < 4; ++c2) {
      float4 o0, o1;
      float *p0 = &o0.x, *p1 = &o1.x;
#pragma unroll
      for (int pos = 0; pos < 4; ++pos) {
        const int t = t0 - 3 + 4 * j + pos;
        const bool in = t >= 0 && t < Tx;
        float e0, e1;
        upk2(a2[pos][c2], e0, e1);
        p0[pos] = in ? elu_f(e0 + xs[2 * c2][4 * j + pos + 2]) : 0.0f;
        p1[pos] = in ? elu_f(e1 + xs[2 * c2 + 1][4 * j + pos + 2]) : 0.0f;
      }
      *reinterpret_cast<float4 *>(&zs[2 * c2][4 * j]) = o0;
      *reinterpret_cast<float4 *>(&zs[2 * c2 + 1][4 * j]) = o1;
    }
  }
  __syncthreads();
  {
    float acc[4];
    const float bfv = bf[0];
#pragma unroll
    for (int pos = 0; pos < 4; ++pos) acc[pos] = bfv;
#pragma unroll
    for (int ci = 0; ci < 8; ++ci) {
      const float4 va = *reinterpret_cast<const float4 *>(&zs[ci][4 * tid]), vb = *reinterpret_cast<const float4 *>(&zs[ci][4 * tid + 4]);
      const float2 vc = *reinterpret_cast<const float2 *>(&zs[ci][4 * tid + 8]);
      const float zin[10] = {va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w, vc.x, vc.y};
#pragma unroll
      for (int k = 0; k < 7; ++k) {
        const float wv = wfs[k][ci];
#pragma unroll
        for (int pos = 0; pos < 4; ++pos) acc[pos] = fmaf(zin[pos + k], wv, acc[pos]);
      }
    }
    const int t = t0 + 4 * tid;
    float *dst = delta + (long long)b * T + t;
    if (4 * tid >= TT) {
    } else if (t + 3 < T && t + 3 < Tx && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
      *reinterpret_cast<float4 *>(dst) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    } else {
#pragma unroll
      for (int pos = 0; pos < 4; ++pos)
        if (t + pos < T) dst[pos] = t + pos < Tx ? acc[pos] : 0.0f;
    }
  }
}

// nn.LSTM(H, H, L <= 2 layers) on channels-first x[b][H][T] (py/main14b_2.py:137,167): one block per clip, 4H threads,
// thread r = gate row r of BOTH layers with its 2 x 2 x H weights in registers; x and y staged through shared memory.
template <int H>
__global__ void __launch_bounds__(4 * H)
    lstm_small_reg_kernel(const float *__restrict__ x, const float *__restrict__ w_ih, const float *__restrict__ w_hh,
                          const float *__restrict__ bias, float *__restrict__ y, int T, int L) {
  constexpr int CH = 64;                       // time steps staged at once
  __shared__ __align__(16) float xin[H], h[2][H], g[4 * H];
  __shared__ float xs[H][CH + 1], ys[H][CH + 1];
  const int b = blockIdx.x, r = threadIdx.x;
  float wi[2][H], wh[2][H], bz[2];
#pragma unroll
  for (int l = 0; l < 2; ++l) {
    const bool on = l < L;
    bz[l] = on ? bias[l * 4 * H + r] : 0.0f;
#pragma unroll
    for (int k = 0; k < H; ++k) {
      wi[l][k] = on ? w_ih[((size_t)l * 4 * H + r) * H + k] : 0.0f;
      wh[l][k] = on ? w_hh[((size_t)l * 4 * H + r) * H + k] : 0.0f;
    }
  }
  float c0 = 0.0f, c1 = 0.0f;                  // cell state of unit r (threads r < H), layers 0 and 1
  if (r < H) { h[0][r] = 0.0f; h[1][r] = 0.0f; }
  for (int tc = 0; tc < T; tc += CH) {
    const int n = T - tc < CH ? T - tc : CH;
    __syncthreads();
    for (int i = r; i < H * n; i += 4 * H) xs[i / n][i % n] = x[((size_t)b * H + i / n) * T + tc + i % n];
    __syncthreads();
    for (int t = 0; t < n; ++t) {
#pragma unroll
      for (int l = 0; l < 2; ++l) {
        if (l < L) {
          if (r < H) xin[r] = l == 0 ? xs[r][t] : h[0][r];
          __syncthreads();
          float a = bz[l], a2 = 0.0f;
#pragma unroll
          for (int k = 0; k < H; k += 4) {
            const float4 xv = *reinterpret_cast<const float4 *>(&xin[k]), hv = *reinterpret_cast<const float4 *>(&h[l][k]);
            a = fmaf(wi[l][k], xv.x, a); a2 = fmaf(wh[l][k], hv.x, a2);
            a = fmaf(wi[l][k + 1], xv.y, a); a2 = fmaf(wh[l][k + 1], hv.y, a2);
            a = fmaf(wi[l][k + 2], xv.z, a); a2 = fmaf(wh[l][k + 2], hv.z, a2);
            a = fmaf(wi[l][k + 3], xv.w, a); a2 = fmaf(wh[l][k + 3], hv.w, a2);
          }
          g[r] = a + a2;
          __syncthreads();
          if (r < H) {
            const float ig = 1.0f / (1.0f + expf(-g[r])), fg = 1.0f / (1.0f + expf(-g[H + r]));
            const float gg = tanhf(g[2 * H + r]), og = 1.0f / (1.0f + expf(-g[3 * H + r]));
            float &cc = l == 0 ? c0 : c1;
            cc = fg * cc + ig * gg;
            h[l][r] = og * tanhf(cc);
          }
          __syncthreads();
        }
      }
      if (r < H) ys[r][t] = h[L - 1][r];
    }
    __syncthreads();
    for (int i = r; i < H * n; i += 4 * H) y[((size_t)b * H + i / n) * T + tc + i % n] = ys[i / n][i % n];
  }
}

}  // namespace

int launch_lstm_small_reg(const float *x, const float *w_ih, const float *w_hh, const float *bias, float *y, int B, int H,
                          int T, int L, cudaStream_t st) {
  if (H != 32 || L < 1 || L > 2) return 1;   // not handled here
  lstm_small_reg_kernel<32><<<B, 128, 0, st>>>(x, w_ih, w_hh, bias, y, T, L);
  WM_CHECK_LAUNCH("lstm_small_reg");
  return 0;
}

}  // namespace wm

using namespace wm;

extern "C" int wm_m14_tail8_fwd(const void *x, long long plane_rows, int B, int Tx, const float *w1, const float *b1,
                                const float *w2, const float *b2, const float *wf, const float *bf, float *delta, int T,
                                void *stream) {
  if (int rc = require_device()) return rc;
  WM_CHECK_ARG(B >= 0 && Tx >= 0 && T >= 0, "m14_tail8: negative size");
  if (B == 0 || T == 0) return 0;
  WM_CHECK_ARG(x && w1 && b1 && w2 && b2 && wf && bf && delta, "m14_tail8: null pointer");
  WM_CHECK_ARG(plane_rows >= wm_pconv_plane_rows(B, Tx), "m14_tail8: plane_rows too small");
  WM_CHECK_ARG(B <= 65535, "m14_tail8: at most 65535 clips per call");
  dim3 grid((T + TT - 1) / TT, B);
  m14_tail8_kernel<<<grid, TN, 0, as_stream(stream)>>>(reinterpret_cast<const uint4 *>(x), plane_rows, Tx, w1, b1, w2, b2,
                                                       wf, bf, delta, T);
  WM_CHECK_LAUNCH("m14_tail8");
  return 0;
}
