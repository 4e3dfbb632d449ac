// nn.LSTM(64,64,batch_first) in training (py/main16.py:138,153): a forward pass that keeps what the backward needs
// (activated gates and cell states of every step) and back-propagation through time.  fp32 CUDA cores, one clip per
// block (training batches are tens of clips; the 16 000-step chain is the cost, not the width).
//
// Weight layout of the training path ("per-gate transposed"): wT[q][k][r] = W[q * 64 + r][k] for gate q of
// (i, f, g, o), which is at once the [ci][co] layout the weight-gradient kernel produces.
//
//   forward  t = 0..T-1   a = b + W_ih x_t + W_hh h_{t-1};  i,f,o = sigmoid, g = tanh;  c_t = f c_{t-1} + i g;
//                         h_t = o tanh(c_t)                              -> gates[q][b][t][64], cell[b][t][64], h
//   backward t = T-1..0   dh = dy_t + W_hh^T da_{t+1};  dc = dc_carry + dh o (1 - tanh^2 c_t);
//                         da = (dc g i(1-i), dc c_{t-1} f(1-f), dc i (1-g^2), dh tanh(c_t) o(1-o));  dc_carry = dc f
//                         -> da[q][b][t][64] (gate-planar: each plane is a dense [rows][64] matrix for the GEMMs)
//   then, parallel over rows: dx = sum_q da_q W_ih,q;  dW_ih,q = x^T da_q;  dW_hh,q = h_{t-1}^T da_q;  db_q = sum da_q
#include "wm_common.h"

namespace wm {

namespace {

__device__ __forceinline__ float sigm(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float tanhx(float x) { return __fdividef(2.0f, 1.0f + __expf(-2.0f * x)) - 1.0f; }

constexpr int TC = 16;   // time steps staged per cp.async chunk

// Packed fp32 FMA (fma.rn.f32x2, SASS FFMA2): two independent IEEE fp32 FMAs per instruction.  The 64-term dots of
// the recurrence are four interleaved partial sums; packing them two by two halves the FMA instructions on the
// step's critical path without changing a bit of the result.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void upk2(f32x2 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ void ffma2(f32x2 &d, f32x2 a, f32x2 b) { asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b)); }

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// b_ih + b_hh
__global__ void bias_sum_kernel(const float *__restrict__ a, const float *__restrict__ b, float *__restrict__ o) {
  o[threadIdx.x] = a[threadIdx.x] + b[threadIdx.x];
}

// The input half of the gates, W_ih x_t + b, does not depend on the recurrence: it is computed for all steps by four
// 1-tap convolutions into the gate planes, and this kernel walks the chain with only the 64-term W_hh h_{t-1} dot
// per thread, overwriting each pre-activation with the activated gate.  thread (q, j) owns gate row q*64 + j.
__global__ void __launch_bounds__(256, 1)
    lstm_train_fwd_kernel(float *__restrict__ gates, const float *__restrict__ wT_hh, float *__restrict__ h_out,
                          float *__restrict__ cell, long long plane, int T) {
  __shared__ __align__(16) float xs[2][TC][256];
  __shared__ __align__(16) float hs[64];
  __shared__ float gs[256];
  const int tid = threadIdx.x, q = tid >> 6, j = tid & 63, b = blockIdx.x;
  f32x2 wh[32];   // (W[k], W[k + 1]) pairs
#pragma unroll
  for (int k = 0; k < 64; k += 2) wh[k >> 1] = pk2(wT_hh[(q * 64 + k) * 64 + j], wT_hh[(q * 64 + k + 1) * 64 + j]);
  float *gq = gates + (size_t)q * plane + (size_t)b * T * 64;
  const float *gb = gates + (size_t)b * T * 64;
  float *hb = h_out + (size_t)b * T * 64, *cb = cell + (size_t)b * T * 64;
  auto stage = [&](int chunk, int buf) {
    const int t0 = chunk * TC;
    for (int i = tid; i < TC * 64; i += 256) {
      const int tt = i >> 6, qq = (i >> 4) & 3, c4 = (i & 15) * 4;
      if (t0 + tt < T) cp_async16(&xs[buf][tt][qq * 64 + c4], gb + (size_t)qq * plane + (size_t)(t0 + tt) * 64 + c4);
    }
    cp_async_commit();
  };
  if (tid < 64) hs[tid] = 0.0f;
  float c_prev = 0.0f;
  const int nchunks = (T + TC - 1) / TC;
  stage(0, 0);
  for (int ch = 0; ch < nchunks; ++ch) {
    const int buf = ch & 1;
    if (ch + 1 < nchunks) { stage(ch + 1, buf ^ 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    __syncthreads();
    const int tend = min(TC, T - ch * TC);
    for (int tt = 0; tt < tend; ++tt) {
      const int t = ch * TC + tt;
      f32x2 a01 = pk2(xs[buf][tt][tid], 0.0f), a23 = pk2(0.0f, 0.0f);
#pragma unroll
      for (int k = 0; k < 64; k += 4) {
        const ulonglong2 hv = *reinterpret_cast<const ulonglong2 *>(&hs[k]);   // (h[k], h[k+1]), (h[k+2], h[k+3])
        ffma2(a01, wh[k >> 1], hv.x);
        ffma2(a23, wh[(k >> 1) + 1], hv.y);
      }
      float a0, a1, a2, a3;
      upk2(a01, a0, a1);
      upk2(a23, a2, a3);
      const float a = (a0 + a1) + (a2 + a3);
      const float g = q == 2 ? tanhx(a) : sigm(a);
      gs[tid] = g;
      gq[(size_t)t * 64 + j] = g;
      __syncthreads();
      if (tid < 64) {
        const float c = fmaf(gs[64 + tid], c_prev, gs[tid] * gs[128 + tid]);
        c_prev = c;
        const float h = gs[192 + tid] * tanhx(c);
        hs[tid] = h;
        hb[(size_t)t * 64 + tid] = h;
        cb[(size_t)t * 64 + tid] = c;
      }
      __syncthreads();
    }
  }
}

// thread (q, j): phase A makes da[q*64 + j] of step t; phase B makes the partial of W_hh^T da over gate q's 64 rows
// for hidden unit j (W_hh[q*64 + rr][j], rr = 0..63, in registers).  Operands of TC steps are staged by cp.async,
// chunks walked from the last to the first.
constexpr int BWD_BUF = TC * 256 + (TC + 1) * 64 + TC * 64;   // floats per staging buffer
__global__ void __launch_bounds__(256, 1)
    lstm_train_bwd_kernel(const float *__restrict__ dy, const float *__restrict__ wT_hh, const float *__restrict__ gates,
                          const float *__restrict__ cell, float *__restrict__ da, long long plane, int T) {
  extern __shared__ __align__(16) float bsm[];
  __shared__ __align__(16) float das[256], part[4][64];
  const int tid = threadIdx.x, q = tid >> 6, j = tid & 63, b = blockIdx.x;
  f32x2 w[32];   // pairs of w[rr] = W_hh[q*64 + rr][j] = wT_hh[q][j][rr]
#pragma unroll
  for (int rr = 0; rr < 64; rr += 4) {
    const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(&wT_hh[(q * 64 + j) * 64 + rr]);
    w[rr >> 1] = v.x; w[(rr >> 1) + 1] = v.y;
  }
  const float *dyb = dy + (size_t)b * T * 64, *gb = gates + (size_t)b * T * 64, *cb = cell + (size_t)b * T * 64;
  float *dab = da + (size_t)q * plane + (size_t)b * T * 64;
  auto stage = [&](int chunk, int buf) {
    float *gsm = bsm + buf * BWD_BUF, *csm = gsm + TC * 256, *dsm = csm + (TC + 1) * 64;
    const int t0 = chunk * TC;
    for (int i = tid; i < TC * 64; i += 256) {
      const int tt = i >> 6, qq = (i >> 4) & 3, c4 = (i & 15) * 4;
      if (t0 + tt < T) cp_async16(&gsm[tt * 256 + qq * 64 + c4], gb + (size_t)qq * plane + (size_t)(t0 + tt) * 64 + c4);
    }
    for (int i = tid; i < (TC + 1) * 16; i += 256) {
      const int r = i >> 4, c4 = (i & 15) * 4, t = t0 - 1 + r;
      if (t >= 0 && t < T) cp_async16(&csm[r * 64 + c4], cb + (size_t)t * 64 + c4);
      else *reinterpret_cast<float4 *>(&csm[r * 64 + c4]) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int i = tid; i < TC * 16; i += 256) {
      const int tt = i >> 4, c4 = (i & 15) * 4;
      if (t0 + tt < T) cp_async16(&dsm[tt * 64 + c4], dyb + (size_t)(t0 + tt) * 64 + c4);
    }
    cp_async_commit();
  };
  part[q][j] = 0.0f;
  float dc_carry = 0.0f;
  const int nchunks = (T + TC - 1) / TC;
  stage(nchunks - 1, 0);
  for (int ci = 0; ci < nchunks; ++ci) {
    const int ch = nchunks - 1 - ci, buf = ci & 1;
    if (ch > 0) { stage(ch - 1, buf ^ 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    __syncthreads();
    const float *gsm = bsm + buf * BWD_BUF, *csm = gsm + TC * 256, *dsm = csm + (TC + 1) * 64;
    const int tend = min(TC, T - ch * TC);
    for (int tt = tend - 1; tt >= 0; --tt) {
      const int t = ch * TC + tt;
      const float gi = gsm[tt * 256 + j], gf = gsm[tt * 256 + 64 + j], gg = gsm[tt * 256 + 128 + j],
                  go = gsm[tt * 256 + 192 + j];
      const float ct = csm[(tt + 1) * 64 + j], cp = csm[tt * 64 + j];
      const float dh = dsm[tt * 64 + j] + (part[0][j] + part[1][j]) + (part[2][j] + part[3][j]);
      const float th = tanhx(ct);
      const float dc = fmaf(dh * go, 1.0f - th * th, dc_carry);
      dc_carry = dc * gf;
      float d;
      if (q == 0) d = dc * gg * gi * (1.0f - gi);
      else if (q == 1) d = dc * cp * gf * (1.0f - gf);
      else if (q == 2) d = dc * gi * (1.0f - gg * gg);
      else d = dh * th * go * (1.0f - go);
      das[tid] = d;
      dab[(size_t)t * 64 + j] = d;
      __syncthreads();
      f32x2 p01 = pk2(0.0f, 0.0f), p23 = p01;
#pragma unroll
      for (int rr = 0; rr < 64; rr += 4) {
        const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(&das[q * 64 + rr]);
        ffma2(p01, w[rr >> 1], v.x);
        ffma2(p23, w[(rr >> 1) + 1], v.y);
      }
      float p0, p1, p2, p3;
      upk2(p01, p0, p1);
      upk2(p23, p2, p3);
      part[q][j] = (p0 + p1) + (p2 + p3);
      __syncthreads();
    }
  }
}

}  // namespace

// x [B][T][64] -> h [B][T][64]; gates [4][B][T][64] (gate-planar, activated) and cell [B][T][64] kept for the backward
int launch_lstm_train_fwd(const float *x, const float *wT_ih, const float *wT_hh, const float *b_ih, const float *b_hh,
                          float *h, float *gates, float *cell, int B, int T, cudaStream_t st) {
  if (B == 0 || T == 0) return 0;
  const size_t plane = (size_t)B * T * 64;
  // gate q's bias sum sits in the first 64 floats of `h` while its convolution runs (h is written later)
  for (int q = 0; q < 4; ++q) {
    bias_sum_kernel<<<1, 64, 0, st>>>(b_ih + q * 64, b_hh + q * 64, h);
    WM_CHECK_LAUNCH("bias_sum");
    WM_TRY(launch_conv64_fp32(x, wT_ih + q * 4096, h, nullptr, nullptr, gates + q * plane, B, T, 1, 0, st));
  }
  lstm_train_fwd_kernel<<<B, 256, 0, st>>>(gates, wT_hh, h, cell, (long long)plane, T);
  WM_CHECK_LAUNCH("lstm_train_fwd");
  return 0;
}

size_t lstm_train_bwd_scratch_floats(int B, int T) {
  const size_t n = (size_t)B * T * 64;
  size_t w = conv_wgrad_scratch_floats(B, T, 1);
  return 4 * n /* da planes */ + n /* dx ping */ + w + 4096 /* transposed gate block */ + 64 /* zero bias */;
}

// dy [B][T][64] -> dx [B][T][64], dwT_ih / dwT_hh [4][64][64], db [256] (the gradient of b_ih and of b_hh)
int launch_lstm_train_bwd(const float *dy, const float *x, const float *h, const float *wT_ih, const float *wT_hh,
                          const float *gates, const float *cell, float *dx, float *dwT_ih, float *dwT_hh, float *db,
                          int B, int T, float *scratch, cudaStream_t st) {
  if (B == 0 || T == 0) return 0;
  const size_t n = (size_t)B * T * 64;
  float *da = scratch, *ping = da + 4 * n, *wg = ping + n, *wt = wg + conv_wgrad_scratch_floats(B, T, 1),
        *zero = wt + 4096;
  constexpr int SMEM = 2 * BWD_BUF * (int)sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    WM_CHECK_CUDA(cudaFuncSetAttribute(lstm_train_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    attr_set = true;
  }
  lstm_train_bwd_kernel<<<B, 256, SMEM, st>>>(dy, wT_hh, gates, cell, da, (long long)n, T);
  WM_CHECK_LAUNCH("lstm_train_bwd");
  WM_CHECK_CUDA(cudaMemsetAsync(zero, 0, 64 * sizeof(float), st));
  for (int q = 0; q < 4; ++q) {
    const float *daq = da + (size_t)q * n;
    WM_TRY(launch_conv_wgrad_ex(x, daq, dwT_ih + q * 4096, db + q * 64, B, T, 1, 0, 0, wg, st));
    WM_TRY(launch_conv_wgrad_ex(h, daq, dwT_hh + q * 4096, nullptr, B, T, 1, 1, -1, wg, st));
    // dx += da_q W_ih,q : as a 1-tap convolution with w[ci = r][co = k] = W_ih[q*64 + r][k] = transpose of wT_ih[q]
    WM_TRY(launch_transpose_flip(wT_ih + q * 4096, wt, 1, st));
    float *out = (q & 1) ? dx : ping;          // ping, dx, ping, dx: the total lands in dx
    const float *acc = q == 0 ? nullptr : ((q & 1) ? ping : dx);
    WM_TRY(launch_conv64_fp32(daq, wt, zero, acc, nullptr, out, B, T, 1, 0, st));
  }
  return 0;
}

}  // namespace wm
