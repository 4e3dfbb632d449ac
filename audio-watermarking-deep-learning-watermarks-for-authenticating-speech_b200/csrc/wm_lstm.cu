// nn.LSTM(64,64,batch_first=True), zero initial state, all hidden states returned
// (py/main16.py:138,153).  fp32 CUDA-core persistent kernel: the 256 gate rows of
// W_ih and W_hh live in the registers of 256 threads for all 16 000 steps; a block
// owns NB clips and walks the time axis once.
//
//   gate[r] = b[r] + W_ih[r,:] . x_t + W_hh[r,:] . h_{t-1}      (thread r, every clip)
//   c_t = sig(f) c_{t-1} + sig(i) tanh(g);  h_t = sig(o) tanh(c_t)   (thread (j, clip))
#include "wm_common.h"

namespace wm {

namespace {
constexpr int kTC = 16;  // time steps of x staged per cp.async chunk

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }
}  // namespace

template <int NB>
__global__ void __launch_bounds__(256, 1)
    lstm_fp32_kernel(const float *__restrict__ x, const float *__restrict__ w_ih,
                     const float *__restrict__ w_hh, const float *__restrict__ bias,
                     float *__restrict__ h_out, int B, int T) {
  extern __shared__ __align__(16) float lstm_smem[];
  float (*xs)[NB][kTC][64] = reinterpret_cast<float (*)[NB][kTC][64]>(lstm_smem);  // [2] staged inputs
  float (*hs)[64] = reinterpret_cast<float (*)[64]>(lstm_smem + 2 * NB * kTC * 64);   // h_{t-1}
  float (*gs)[256] = reinterpret_cast<float (*)[256]>(lstm_smem + 2 * NB * kTC * 64 + NB * 64);  // gates
  const int tid = threadIdx.x;
  const int b0 = blockIdx.x * NB;
  const int nb = min(NB, B - b0);

  float wi[64], wh[64];
#pragma unroll
  for (int k = 0; k < 64; k += 4) {
    float4 a = *reinterpret_cast<const float4 *>(&w_ih[tid * 64 + k]);
    float4 c = *reinterpret_cast<const float4 *>(&w_hh[tid * 64 + k]);
    wi[k] = a.x; wi[k + 1] = a.y; wi[k + 2] = a.z; wi[k + 3] = a.w;
    wh[k] = c.x; wh[k + 1] = c.y; wh[k + 2] = c.z; wh[k + 3] = c.w;
  }
  const float br = bias[tid];
  const bool is_tanh = (tid >= 128 && tid < 192);  // rows i,f,g,o: only g uses tanh (warp-uniform)

  for (int i = tid; i < NB * 64; i += 256) (&hs[0][0])[i] = 0.0f;
  // cell state: thread owns unit j = tid % 64 of clips n = tid / 64 + 4 m
  constexpr int NM = (NB + 3) / 4;
  float c_state[NM];
#pragma unroll
  for (int m = 0; m < NM; ++m) c_state[m] = 0.0f;
  const int uj = tid & 63, un = tid >> 6;

  auto stage = [&](int chunk, int buf) {
    // NB clips x kTC steps x 64 floats -> 16-byte cp.async, (NB * kTC * 16) of them
    int tbase = chunk * kTC;
    for (int i = tid; i < NB * kTC * 16; i += 256) {
      int n = i / (kTC * 16), rem = i % (kTC * 16), tt = rem >> 4, c4 = (rem & 15) * 4;
      int t = tbase + tt;
      if (n < nb && t < T) cp_async16(&xs[buf][n][tt][c4], &x[((size_t)(b0 + n) * T + t) * 64 + c4]);
    }
    cp_async_commit();
  };

  const int nchunks = (T + kTC - 1) / kTC;
  stage(0, 0);
  for (int ch = 0; ch < nchunks; ++ch) {
    const int buf = ch & 1;
    if (ch + 1 < nchunks) { stage(ch + 1, buf ^ 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    __syncthreads();
    const int tend = min(kTC, T - ch * kTC);
    for (int tt = 0; tt < tend; ++tt) {
      float acc[NB];
#pragma unroll
      for (int n = 0; n < NB; ++n) acc[n] = br;
#pragma unroll
      for (int k = 0; k < 64; k += 4) {
#pragma unroll
        for (int n = 0; n < NB; ++n) {
          float4 xv = *reinterpret_cast<const float4 *>(&xs[buf][n][tt][k]);
          float4 hv = *reinterpret_cast<const float4 *>(&hs[n][k]);
          acc[n] = fmaf(wi[k], xv.x, acc[n]); acc[n] = fmaf(wh[k], hv.x, acc[n]);
          acc[n] = fmaf(wi[k + 1], xv.y, acc[n]); acc[n] = fmaf(wh[k + 1], hv.y, acc[n]);
          acc[n] = fmaf(wi[k + 2], xv.z, acc[n]); acc[n] = fmaf(wh[k + 2], hv.z, acc[n]);
          acc[n] = fmaf(wi[k + 3], xv.w, acc[n]); acc[n] = fmaf(wh[k + 3], hv.w, acc[n]);
        }
      }
#pragma unroll
      for (int n = 0; n < NB; ++n) gs[n][tid] = is_tanh ? tanh_acc(acc[n]) : sigmoid_acc(acc[n]);
      __syncthreads();
      const int t = ch * kTC + tt;
#pragma unroll
      for (int m = 0; m < NM; ++m) {
        int n = un + 4 * m;
        if (n < NB) {
          float ig = gs[n][uj], fg = gs[n][64 + uj], gg = gs[n][128 + uj], og = gs[n][192 + uj];
          float c = fmaf(fg, c_state[m], ig * gg);
          c_state[m] = c;
          float h = og * tanh_acc(c);
          hs[n][uj] = h;
          if (n < nb) h_out[((size_t)(b0 + n) * T + t) * 64 + uj] = h;
        }
      }
      __syncthreads();
    }
  }
}

template <int NB>
static int launch_lstm_nb(const float *x, const float *w_ih, const float *w_hh, const float *bias,
                          float *h, int B, int T, cudaStream_t st) {
  constexpr int SMEM = (2 * NB * kTC * 64 + NB * 64 + NB * 256) * 4;
  static bool attr_set = false;
  if (!attr_set) {
    WM_CHECK_CUDA(cudaFuncSetAttribute(lstm_fp32_kernel<NB>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    attr_set = true;
  }
  lstm_fp32_kernel<NB><<<(B + NB - 1) / NB, 256, SMEM, st>>>(x, w_ih, w_hh, bias, h, B, T);
  WM_CHECK_LAUNCH("lstm_fp32");
  return 0;
}

int launch_lstm_fp32(const float *x, const float *w_ih, const float *w_hh, const float *bias,
                     float *h, int B, int T, cudaStream_t st) {
  if (B == 0 || T == 0) return 0;
  // clips per block: spread B over a whole number of waves of one block per SM, i.e. the
  // smallest NB <= 8 with ceil(B / NB) <= k * SMs for the smallest possible k
  int sms = sm_count();
  int k = (B + sms * 8 - 1) / (sms * 8);
  int nb = (B + k * sms - 1) / (k * sms);
  switch (nb) {
    case 1: return launch_lstm_nb<1>(x, w_ih, w_hh, bias, h, B, T, st);
    case 2: return launch_lstm_nb<2>(x, w_ih, w_hh, bias, h, B, T, st);
    case 3: return launch_lstm_nb<3>(x, w_ih, w_hh, bias, h, B, T, st);
    case 4: return launch_lstm_nb<4>(x, w_ih, w_hh, bias, h, B, T, st);
    case 5: return launch_lstm_nb<5>(x, w_ih, w_hh, bias, h, B, T, st);
    case 6: return launch_lstm_nb<6>(x, w_ih, w_hh, bias, h, B, T, st);
    case 7: return launch_lstm_nb<7>(x, w_ih, w_hh, bias, h, B, T, st);
    default: return launch_lstm_nb<8>(x, w_ih, w_hh, bias, h, B, T, st);
  }
}

}  // namespace wm
