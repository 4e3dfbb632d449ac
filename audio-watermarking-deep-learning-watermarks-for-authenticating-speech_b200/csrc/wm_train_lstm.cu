// nn.LSTM(64,64,batch_first) in training (py/main16.py:138,153): a forward pass that keeps what the backward needs
// (activated gates and cell states of every step) and back-propagation through time.  fp32 CUDA cores, one clip per
// block (training batches are tens of clips; the 16 000-step chain is the cost, not the width).
//
// Weight layout of the training path ("per-gate transposed"): wT[q][k][r] = W[q * 64 + r][k] for gate q of
// (i, f, g, o), which is at once the [ci][co] layout the weight-gradient kernel produces.
//
//   forward  t = 0..T-1   a = b + W_ih x_t + W_hh h_{t-1};  i,f,o = sigmoid, g = tanh;  c_t = f c_{t-1} + i g;
//                         h_t = o tanh(c_t)                              -> gates[b][t][4][64], cell[b][t][64], h
//   backward t = T-1..0   dh = dy_t + W_hh^T da_{t+1};  dc = dc_carry + dh o (1 - tanh^2 c_t);
//                         da = (dc g i(1-i), dc c_{t-1} f(1-f), dc i (1-g^2), dh tanh(c_t) o(1-o));  dc_carry = dc f
//                         -> da[q][b][t][64] (gate-planar: each plane is a dense [rows][64] matrix for the GEMMs)
//   then, parallel over rows: dx = sum_q da_q W_ih,q;  dW_ih,q = x^T da_q;  dW_hh,q = h_{t-1}^T da_q;  db_q = sum da_q
#include "wm_common.h"

namespace wm {

namespace {

__device__ __forceinline__ float sigm(float x) { return 1.0f / (1.0f + __expf(-x)); }
__device__ __forceinline__ float tanhx(float x) { return 2.0f / (1.0f + __expf(-2.0f * x)) - 1.0f; }

// thread r = q * 64 + j owns gate row r (W_ih[r,:], W_hh[r,:] in registers)
__global__ void __launch_bounds__(256, 1)
    lstm_train_fwd_kernel(const float *__restrict__ x, const float *__restrict__ wT_ih, const float *__restrict__ wT_hh,
                          const float *__restrict__ b_ih, const float *__restrict__ b_hh, float *__restrict__ h_out,
                          float *__restrict__ gates, float *__restrict__ cell, int T) {
  __shared__ __align__(16) float xs[2][64], hs[64], gs[256];
  const int tid = threadIdx.x, q = tid >> 6, j = tid & 63, b = blockIdx.x;
  float wi[64], wh[64];
#pragma unroll
  for (int k = 0; k < 64; ++k) {
    wi[k] = wT_ih[(q * 64 + k) * 64 + j];
    wh[k] = wT_hh[(q * 64 + k) * 64 + j];
  }
  const float br = b_ih[tid] + b_hh[tid];
  const float *xb = x + (size_t)b * T * 64;
  float *hb = h_out + (size_t)b * T * 64, *gb = gates + (size_t)b * T * 256, *cb = cell + (size_t)b * T * 64;
  if (tid < 64) { hs[tid] = 0.0f; xs[0][tid] = xb[tid]; }
  float c_prev = 0.0f;
  __syncthreads();
  for (int t = 0; t < T; ++t) {
    const int buf = t & 1;
    float xn = 0.0f;
    if (tid < 64 && t + 1 < T) xn = xb[(size_t)(t + 1) * 64 + tid];   // next step's input, in flight during the dot
    float a0 = br, a1 = 0.0f;
#pragma unroll
    for (int k = 0; k < 64; k += 4) {
      const float4 xv = *reinterpret_cast<const float4 *>(&xs[buf][k]);
      const float4 hv = *reinterpret_cast<const float4 *>(&hs[k]);
      a0 = fmaf(wi[k], xv.x, a0); a1 = fmaf(wh[k], hv.x, a1);
      a0 = fmaf(wi[k + 1], xv.y, a0); a1 = fmaf(wh[k + 1], hv.y, a1);
      a0 = fmaf(wi[k + 2], xv.z, a0); a1 = fmaf(wh[k + 2], hv.z, a1);
      a0 = fmaf(wi[k + 3], xv.w, a0); a1 = fmaf(wh[k + 3], hv.w, a1);
    }
    const float a = a0 + a1;
    const float g = q == 2 ? tanhx(a) : sigm(a);
    gs[tid] = g;
    gb[(size_t)t * 256 + tid] = g;
    __syncthreads();
    if (tid < 64) {
      const float c = fmaf(gs[64 + tid], c_prev, gs[tid] * gs[128 + tid]);
      c_prev = c;
      const float h = gs[192 + tid] * tanhx(c);
      hs[tid] = h;
      hb[(size_t)t * 64 + tid] = h;
      cb[(size_t)t * 64 + tid] = c;
      xs[buf ^ 1][tid] = xn;
    }
    __syncthreads();
  }
}

// thread (q, j): phase A makes da[q*64 + j] of step t; phase B makes the partial of W_hh^T da over gate q's 64 rows
// for hidden unit j (W_hh[q*64 + rr][j], rr = 0..63, in registers).
__global__ void __launch_bounds__(256, 1)
    lstm_train_bwd_kernel(const float *__restrict__ dy, const float *__restrict__ wT_hh, const float *__restrict__ gates,
                          const float *__restrict__ cell, float *__restrict__ da, long long plane, int T) {
  __shared__ __align__(16) float das[256], part[4][64];
  const int tid = threadIdx.x, q = tid >> 6, j = tid & 63, b = blockIdx.x;
  float w[64];   // w[rr] = W_hh[q*64 + rr][j] = wT_hh[q][j][rr]
#pragma unroll
  for (int rr = 0; rr < 64; rr += 4) {
    const float4 v = *reinterpret_cast<const float4 *>(&wT_hh[(q * 64 + j) * 64 + rr]);
    w[rr] = v.x; w[rr + 1] = v.y; w[rr + 2] = v.z; w[rr + 3] = v.w;
  }
  const float *dyb = dy + (size_t)b * T * 64, *gb = gates + (size_t)b * T * 256, *cb = cell + (size_t)b * T * 64;
  float *dab = da + (size_t)q * plane + (size_t)b * T * 64;
  part[q][j] = 0.0f;
  float dc_carry = 0.0f;
  // operands of step t, loaded one step ahead
  float gi, gf, gg, go, ct, cp, dyt;
  auto load = [&](int t, float &i_, float &f_, float &g_, float &o_, float &c_, float &cprev_, float &dy_) {
    const float *gr = gb + (size_t)t * 256;
    i_ = gr[j]; f_ = gr[64 + j]; g_ = gr[128 + j]; o_ = gr[192 + j];
    c_ = cb[(size_t)t * 64 + j];
    cprev_ = t > 0 ? cb[(size_t)(t - 1) * 64 + j] : 0.0f;
    dy_ = dyb[(size_t)t * 64 + j];
  };
  load(T - 1, gi, gf, gg, go, ct, cp, dyt);
  __syncthreads();
  for (int t = T - 1; t >= 0; --t) {
    float ni = 0, nf = 0, ng = 0, no = 0, nc = 0, ncp = 0, ndy = 0;
    if (t > 0) load(t - 1, ni, nf, ng, no, nc, ncp, ndy);
    const float dh = dyt + (part[0][j] + part[1][j]) + (part[2][j] + part[3][j]);
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
    float p0 = 0.0f, p1 = 0.0f;
#pragma unroll
    for (int rr = 0; rr < 64; rr += 4) {
      const float4 v = *reinterpret_cast<const float4 *>(&das[q * 64 + rr]);
      p0 = fmaf(w[rr], v.x, p0); p1 = fmaf(w[rr + 1], v.y, p1);
      p0 = fmaf(w[rr + 2], v.z, p0); p1 = fmaf(w[rr + 3], v.w, p1);
    }
    part[q][j] = p0 + p1;
    gi = ni; gf = nf; gg = ng; go = no; ct = nc; cp = ncp; dyt = ndy;
    __syncthreads();
  }
}

}  // namespace

// x [B][T][64] -> h [B][T][64]; gates [B][T][256], cell [B][T][64] kept for the backward
int launch_lstm_train_fwd(const float *x, const float *wT_ih, const float *wT_hh, const float *b_ih, const float *b_hh,
                          float *h, float *gates, float *cell, int B, int T, cudaStream_t st) {
  if (B == 0 || T == 0) return 0;
  lstm_train_fwd_kernel<<<B, 256, 0, st>>>(x, wT_ih, wT_hh, b_ih, b_hh, h, gates, cell, T);
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
  lstm_train_bwd_kernel<<<B, 256, 0, st>>>(dy, wT_hh, gates, cell, da, (long long)n, T);
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
