// Training-loss forward kernels of main16 (py/main16.py:74-81, 192-217, 255-266):
//   * torch.stft(x, N, hop, hann_window(N)) magnitude (centre + reflect padding, periodic Hann, onesided)
//   * high_freq_penalty       mean over (B, N/2+1, F) of |STFT_512/128(delta)| * [bin > first_bin]
//   * TFLoudnessLoss          mean of (|S_w| - |S_c|)^2 * [|S_c| > 0.01], STFT 2048/512
//   * MultiScaleMelLoss       mean |log(mel(s)+1e-5) - log(mel(s_w)+1e-5)|, mel = |STFT_1024/256|^2 x fb(513x64)
//   * BCE-with-logits of the detection channel and of the message-bit channels, mean |delta|
//
// One staged FFT serves all of them: a CTA owns one *pair* of real frames, packs them as the real and
// imaginary part of one complex signal (clean + i*watermarked for the two-signal losses, frames f and f+1
// for the one-signal ones), runs log2(N) radix-2 stages in shared memory (bit-reversed load, twiddles from
// a table built once per CTA with sincospif) and separates the two spectra by conjugate symmetry.  The
// loss-specific reduction happens on the spectrum while it is still in shared memory; per-CTA partial sums
// go to the workspace and a single-block kernel adds them in a fixed order (deterministic results).
// These kernels are HBM-trivial (each sample is read ~4x from L2) and ~8 MFLOP per clip: < 0.1 % of a
// training step, so they stay on the CUDA cores.
#include "wm_common.h"

namespace wm {

namespace {

constexpr int FFT_THREADS = 256;
enum { MODE_MAG = 0, MODE_HF = 1, MODE_LOUD = 2, MODE_MEL = 3 };

__device__ __forceinline__ int reflect_idx(int j, int T) {
  // torch 'reflect' padding (no edge repeat); valid while the pad is < T
  if (j < 0) j = -j;
  if (j >= T) j = 2 * (T - 1) - j;
  return j;
}

__device__ __forceinline__ float block_sum(float v, float *red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.0f;
  if (threadIdx.x == 0) {
    for (int i = 0; i < FFT_THREADS / 32; ++i) t += red[i];
  }
  return t;   // valid on thread 0
}

// z[N] holds the bit-reversed input; in-place decimation-in-time radix-2 stages
template <int N, int LOG2N>
__device__ __forceinline__ void fft_stages(float2 *z, const float2 *tw) {
#pragma unroll 1
  for (int s = 0; s < LOG2N; ++s) {
    const int half = 1 << s;
    __syncthreads();
    for (int j = threadIdx.x; j < N / 2; j += FFT_THREADS) {
      const int pos = j & (half - 1);
      const int i0 = ((j >> s) << (s + 1)) + pos, i1 = i0 + half;
      const float2 w = tw[pos << (LOG2N - 1 - s)];
      const float2 a = z[i0], b = z[i1];
      const float2 t = make_float2(w.x * b.x - w.y * b.y, w.x * b.y + w.y * b.x);
      z[i0] = make_float2(a.x + t.x, a.y + t.y);
      z[i1] = make_float2(a.x - t.x, a.y - t.y);
    }
  }
  __syncthreads();
}

// spectra of the two packed real signals at bin k (0 <= k <= N/2)
template <int N>
__device__ __forceinline__ void unpack_pair(const float2 *z, int k, float2 &A, float2 &Bv) {
  const float2 zk = z[k], zn = z[(N - k) & (N - 1)];
  A = make_float2(0.5f * (zk.x + zn.x), 0.5f * (zk.y - zn.y));
  Bv = make_float2(0.5f * (zk.y + zn.y), -0.5f * (zk.x - zn.x));
}

// items: MODE_MAG / MODE_HF  (clip b, frame pair fp) of x0;   MODE_LOUD / MODE_MEL  (clip b, frame f) of (x0, x1)
template <int N, int LOG2N, int MODE>
__global__ void __launch_bounds__(FFT_THREADS)
    stft_kernel(const float *__restrict__ x0, const float *__restrict__ x1, int B, int T, int hop, int F,
                float *__restrict__ mag, float *__restrict__ partials, int first_bin, float thresh,
                const float *__restrict__ fb, const int *__restrict__ band, int n_mels) {
  __shared__ float2 z[N];
  __shared__ float2 tw[N / 2];
  __shared__ float win[N];
  __shared__ float red[FFT_THREADS / 32];
  constexpr bool TWO_SIGNALS = MODE == MODE_LOUD || MODE == MODE_MEL;
  __shared__ float pw[MODE == MODE_MEL ? 2 * (N / 2 + 1) : 2];
  const int per_clip = TWO_SIGNALS ? F : (F + 1) / 2;
  const long long items = (long long)B * per_clip;

  for (int k = threadIdx.x; k < N / 2; k += FFT_THREADS) {
    float sn, cs;
    sincospif(-2.0f * (float)k / (float)N, &sn, &cs);
    tw[k] = make_float2(cs, sn);
  }
  for (int n = threadIdx.x; n < N; n += FFT_THREADS) {
    float sn = sinpif((float)n / (float)N);          // periodic Hann: 0.5 - 0.5 cos(2 pi n / N) = sin^2(pi n / N)
    win[n] = sn * sn;
  }
  __syncthreads();

  for (long long item = blockIdx.x; item < items; item += gridDim.x) {
    const int b = (int)(item / per_clip), u = (int)(item % per_clip);
    const int fa = TWO_SIGNALS ? u : 2 * u;           // frame of the real part
    const int fbm = TWO_SIGNALS ? u : 2 * u + 1;      // frame of the imaginary part
    const bool has_b = TWO_SIGNALS || fbm < F;
    const float *pa = x0 + (size_t)b * T;
    const float *pb = (TWO_SIGNALS ? x1 : x0) + (size_t)b * T;
    __syncthreads();                                   // previous item's readers are done with z
    for (int n = threadIdx.x; n < N; n += FFT_THREADS) {
      const int ja = reflect_idx(fa * hop + n - N / 2, T);
      const int jb = reflect_idx(fbm * hop + n - N / 2, T);
      const float w = win[n];
      const float va = __ldg(pa + ja) * w;
      const float vb = has_b ? __ldg(pb + jb) * w : 0.0f;
      z[__brev((unsigned)n) >> (32 - LOG2N)] = make_float2(va, vb);
    }
    fft_stages<N, LOG2N>(z, tw);

    if constexpr (MODE == MODE_MAG) {
      for (int k = threadIdx.x; k <= N / 2; k += FFT_THREADS) {
        float2 A, Bv;
        unpack_pair<N>(z, k, A, Bv);
        float *dst = mag + ((size_t)b * (N / 2 + 1) + k) * F;
        dst[fa] = sqrtf(A.x * A.x + A.y * A.y);
        if (has_b) dst[fbm] = sqrtf(Bv.x * Bv.x + Bv.y * Bv.y);
      }
    } else if constexpr (MODE == MODE_HF) {
      float acc = 0.0f;
      for (int k = first_bin + threadIdx.x; k <= N / 2; k += FFT_THREADS) {
        float2 A, Bv;
        unpack_pair<N>(z, k, A, Bv);
        acc += sqrtf(A.x * A.x + A.y * A.y);
        if (has_b) acc += sqrtf(Bv.x * Bv.x + Bv.y * Bv.y);
      }
      const float t = block_sum(acc, red);
      if (threadIdx.x == 0) partials[item] = t;
    } else if constexpr (MODE == MODE_LOUD) {
      float acc = 0.0f;
      for (int k = threadIdx.x; k <= N / 2; k += FFT_THREADS) {
        float2 C, Wv;
        unpack_pair<N>(z, k, C, Wv);
        const float mc = sqrtf(C.x * C.x + C.y * C.y), mw = sqrtf(Wv.x * Wv.x + Wv.y * Wv.y);
        const float d = mw - mc;
        acc += mc > thresh ? d * d : 0.0f;
      }
      const float t = block_sum(acc, red);
      if (threadIdx.x == 0) partials[item] = t;
    } else {
      constexpr int NB = N / 2 + 1;
      for (int k = threadIdx.x; k < NB; k += FFT_THREADS) {
        float2 C, Wv;
        unpack_pair<N>(z, k, C, Wv);
        pw[k] = C.x * C.x + C.y * C.y;
        pw[NB + k] = Wv.x * Wv.x + Wv.y * Wv.y;
      }
      __syncthreads();
      float acc = 0.0f;
      for (int m = threadIdx.x; m < n_mels; m += FFT_THREADS) {
        // banded mel projection: filter m is non-zero on bins band[2m] .. band[2m+1]-1
        float mc = 0.0f, mw = 0.0f;
        for (int k = band[2 * m]; k < band[2 * m + 1]; ++k) {
          const float f = __ldg(fb + (size_t)k * n_mels + m);
          mc = fmaf(f, pw[k], mc);
          mw = fmaf(f, pw[NB + k], mw);
        }
        acc += fabsf(logf(mc + 1e-5f) - logf(mw + 1e-5f));
      }
      const float t = block_sum(acc, red);
      if (threadIdx.x == 0) partials[item] = t;
    }
  }
}

// out[0] = scale * sum(partials[0..n))   — one block, double accumulation, fixed order
__global__ void __launch_bounds__(1024) sum_partials_kernel(const float *__restrict__ partials, long long n,
                                                            double scale, float *__restrict__ out) {
  __shared__ double red[1024];
  double a = 0.0;
  for (long long i = threadIdx.x; i < n; i += 1024) a += (double)partials[i];
  red[threadIdx.x] = a;
  __syncthreads();
  for (int s = 512; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = (float)(red[0] * scale);
}

// BCE-with-logits (py/main16.py:255-264): logits [B2][T][nout]; channel 0 against 1 for clips < B_wm else 0,
// channels 1.. of clips < B_wm against bit j of message[b].  partials: [2][gridDim.x]
__global__ void __launch_bounds__(FFT_THREADS)
    bce_heads_kernel(const float *__restrict__ logits, const long long *__restrict__ message, int B_wm, int B2, int T,
                     int nout, float *__restrict__ partials) {
  __shared__ float red[FFT_THREADS / 32];
  const long long rows = (long long)B2 * T;
  float loc = 0.0f, bit = 0.0f;
  for (long long r = (long long)blockIdx.x * FFT_THREADS + threadIdx.x; r < rows; r += (long long)gridDim.x * FFT_THREADS) {
    const int b = (int)(r / T);
    const float *p = logits + r * nout;
    const float x = p[0];
    const float y = b < B_wm ? 1.0f : 0.0f;
    loc += fmaxf(x, 0.0f) - x * y + log1pf(expf(-fabsf(x)));
    if (b < B_wm && nout > 1) {
      const long long m = message[b];
      for (int j = 1; j < nout; ++j) {
        const float v = p[j];
        const float yb = (float)((m >> (j - 1)) & 1);
        bit += fmaxf(v, 0.0f) - v * yb + log1pf(expf(-fabsf(v)));
      }
    }
  }
  const float t0 = block_sum(loc, red);
  if (threadIdx.x == 0) partials[blockIdx.x] = t0;
  const float t1 = block_sum(bit, red);
  if (threadIdx.x == 0) partials[gridDim.x + blockIdx.x] = t1;
}

__global__ void __launch_bounds__(FFT_THREADS)
    abs_sum_kernel(const float *__restrict__ x, long long n, float *__restrict__ partials) {
  __shared__ float red[FFT_THREADS / 32];
  float a = 0.0f;
  for (long long i = (long long)blockIdx.x * FFT_THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * FFT_THREADS)
    a += fabsf(x[i]);
  const float t = block_sum(a, red);
  if (threadIdx.x == 0) partials[blockIdx.x] = t;
}

template <int MODE>
int launch_stft(const float *x0, const float *x1, int B, int T, int n_fft, int hop, float *mag, float *partials,
                int first_bin, float thresh, const float *fb, const int *band, int n_mels, int *items_out,
                cudaStream_t st) {
  const int F = 1 + T / hop;
  const bool two = MODE == MODE_LOUD || MODE == MODE_MEL;
  const long long items = (long long)B * (two ? F : (F + 1) / 2);
  if (items_out) *items_out = (int)items;
  if (items == 0) return 0;
  if (T <= n_fft / 2) {
    set_error("stft: reflect padding needs T > n_fft/2 (T=%d, n_fft=%d)", T, n_fft);
    return -1;
  }
  const int grid = (int)(items < 8LL * sm_count() ? items : 8LL * sm_count());
  switch (n_fft) {
    case 512:
      stft_kernel<512, 9, MODE><<<grid, FFT_THREADS, 0, st>>>(x0, x1, B, T, hop, F, mag, partials, first_bin, thresh, fb, band, n_mels);
      break;
    case 1024:
      stft_kernel<1024, 10, MODE><<<grid, FFT_THREADS, 0, st>>>(x0, x1, B, T, hop, F, mag, partials, first_bin, thresh, fb, band, n_mels);
      break;
    case 2048:
      stft_kernel<2048, 11, MODE><<<grid, FFT_THREADS, 0, st>>>(x0, x1, B, T, hop, F, mag, partials, first_bin, thresh, fb, band, n_mels);
      break;
    default:
      set_error("stft: n_fft must be 512, 1024 or 2048 (got %d)", n_fft);
      return -1;
  }
  WM_CHECK_LAUNCH("stft");
  return 0;
}

// ---- backward of the STFT losses ---------------------------------------------------------------------------
// For a real frame x[n] and its one-sided spectrum X[k] (k <= N/2), with G[k] = dL/dRe X[k] + i dL/dIm X[k]:
//   dL/dx[n] = Re sum_{k <= N/2} G[k] e^{+2 pi i k n / N} = Re FFT(conj G)[n]
// so the same forward stages turn the spectrum gradient into the frame gradient.  A CTA recomputes the forward
// spectra of its item (as stft_kernel does), forms G of the signal the gradient is wanted for, transforms it and
// writes win[n] * Re(.) to gframes[b][f][N]; overlap_add_kernel then GATHERS, per sample, the frames (and the
// reflected padding positions) that touch it — no atomics, so the gradient is deterministic.
//   MODE_HF:   x0 = delta, gradient w.r.t. delta     G = coef * [k >= first_bin] * S / |S|
//   MODE_LOUD: gradient w.r.t. x1 (watermarked)      G = coef * 2 [ |C| > thresh ] (|W| - |C|) W / |W|
//   MODE_MEL:  gradient w.r.t. x1                    G = 2 W sum_m fb[k][m] dmel[m],
//                                                    dmel[m] = -coef * sign(log(mc+eps) - log(mw+eps)) / (mw+eps)
// coef carries 1 / (number of terms of the mean).
template <int N, int LOG2N, int MODE>
__global__ void __launch_bounds__(FFT_THREADS)
    stft_bwd_kernel(const float *__restrict__ x0, const float *__restrict__ x1, int B, int T, int hop, int F,
                    float *__restrict__ gframes, int first_bin, float thresh, const float *__restrict__ fb,
                    const int *__restrict__ band, int n_mels, float coef) {
  __shared__ float2 z[N];
  __shared__ float2 zg[N];
  __shared__ float2 tw[N / 2];
  constexpr bool TWO_SIGNALS = MODE == MODE_LOUD || MODE == MODE_MEL;
  constexpr int NB = N / 2 + 1;
  __shared__ float pw[MODE == MODE_MEL ? 2 * NB : 2];
  __shared__ float dmel[MODE == MODE_MEL ? 128 : 2];
  const int per_clip = TWO_SIGNALS ? F : (F + 1) / 2;
  const long long items = (long long)B * per_clip;
  auto winf = [](int n) { const float sn = sinpif((float)n / (float)N); return sn * sn; };

  for (int k = threadIdx.x; k < N / 2; k += FFT_THREADS) {
    float sn, cs;
    sincospif(-2.0f * (float)k / (float)N, &sn, &cs);
    tw[k] = make_float2(cs, sn);
  }
  __syncthreads();

  for (long long item = blockIdx.x; item < items; item += gridDim.x) {
    const int b = (int)(item / per_clip), u = (int)(item % per_clip);
    const int fa = TWO_SIGNALS ? u : 2 * u;
    const int fbm = TWO_SIGNALS ? u : 2 * u + 1;
    const bool has_b = TWO_SIGNALS || fbm < F;
    const float *pa = x0 + (size_t)b * T;
    const float *pb = (TWO_SIGNALS ? x1 : x0) + (size_t)b * T;
    __syncthreads();
    for (int n = threadIdx.x; n < N; n += FFT_THREADS) {
      const int ja = reflect_idx(fa * hop + n - N / 2, T);
      const int jb = reflect_idx(fbm * hop + n - N / 2, T);
      const float w = winf(n);
      const float va = __ldg(pa + ja) * w;
      const float vb = has_b ? __ldg(pb + jb) * w : 0.0f;
      z[__brev((unsigned)n) >> (32 - LOG2N)] = make_float2(va, vb);
    }
    fft_stages<N, LOG2N>(z, tw);

    if constexpr (MODE == MODE_MEL) {
      for (int k = threadIdx.x; k < NB; k += FFT_THREADS) {
        float2 C, Wv;
        unpack_pair<N>(z, k, C, Wv);
        pw[k] = C.x * C.x + C.y * C.y;
        pw[NB + k] = Wv.x * Wv.x + Wv.y * Wv.y;
      }
      __syncthreads();
      for (int m = threadIdx.x; m < n_mels; m += FFT_THREADS) {
        float mc = 0.0f, mw = 0.0f;
        for (int k = band[2 * m]; k < band[2 * m + 1]; ++k) {
          const float f = __ldg(fb + (size_t)k * n_mels + m);
          mc = fmaf(f, pw[k], mc);
          mw = fmaf(f, pw[NB + k], mw);
        }
        const float d = logf(mc + 1e-5f) - logf(mw + 1e-5f);
        dmel[m] = -coef * (float)((d > 0.0f) - (d < 0.0f)) / (mw + 1e-5f);
      }
      __syncthreads();
    }

    // one inverse transform per target frame: HF has two (a, b), the two-signal losses one (the watermarked frame)
    const int ntargets = TWO_SIGNALS ? 1 : (has_b ? 2 : 1);
    for (int tg = 0; tg < ntargets; ++tg) {
      for (int k = threadIdx.x; k < N; k += FFT_THREADS) {
        float2 g = make_float2(0.0f, 0.0f);
        if (k < NB) {
          float2 A, Bv;
          unpack_pair<N>(z, k, A, Bv);
          if constexpr (MODE == MODE_HF) {
            const float2 S = tg == 0 ? A : Bv;
            const float m = sqrtf(S.x * S.x + S.y * S.y);
            if (k >= first_bin && m > 0.0f) g = make_float2(coef * S.x / m, coef * S.y / m);
          } else if constexpr (MODE == MODE_LOUD) {
            const float mc = sqrtf(A.x * A.x + A.y * A.y), mw = sqrtf(Bv.x * Bv.x + Bv.y * Bv.y);
            if (mc > thresh && mw > 0.0f) {
              const float c = coef * 2.0f * (mw - mc) / mw;
              g = make_float2(c * Bv.x, c * Bv.y);
            }
          } else {
            float gs = 0.0f;
            for (int m = 0; m < n_mels; ++m)
              if (k >= band[2 * m] && k < band[2 * m + 1]) gs = fmaf(__ldg(fb + (size_t)k * n_mels + m), dmel[m], gs);
            g = make_float2(2.0f * gs * Bv.x, 2.0f * gs * Bv.y);
          }
        }
        zg[__brev((unsigned)k) >> (32 - LOG2N)] = make_float2(g.x, -g.y);   // conj(G)
      }
      fft_stages<N, LOG2N>(zg, tw);
      const int f = TWO_SIGNALS ? fa : (tg == 0 ? fa : fbm);
      float *dst = gframes + ((size_t)b * F + f) * N;
      for (int n = threadIdx.x; n < N; n += FFT_THREADS) dst[n] = winf(n) * zg[n].x;
      __syncthreads();
    }
  }
}

// dx[b][t] (+)= sum of gframes over every (frame, offset) whose padded position maps to sample t
__global__ void __launch_bounds__(256)
    overlap_add_kernel(const float *__restrict__ gframes, float *__restrict__ dx, int T, int N, int hop, int F,
                       int accumulate) {
  const int b = blockIdx.y, t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const float *g = gframes + (size_t)b * F * N;
  float acc = 0.0f;
  auto add_pos = [&](int p) {   // p: index into the reflect-padded signal of length T + N
    int fhi = p / hop;
    if (fhi > F - 1) fhi = F - 1;
    int flo = p - N + 1 <= 0 ? 0 : (p - N + hop) / hop;   // ceil((p - N + 1) / hop)
    for (int f = flo; f <= fhi; ++f) acc += g[(size_t)f * N + (p - f * hop)];
  };
  add_pos(t + N / 2);
  if (t >= 1 && t <= N / 2) add_pos(N / 2 - t);
  const int p2 = 2 * (T - 1) - t + N / 2;
  if (t <= T - 2 && p2 < T + N) add_pos(p2);
  float *o = dx + (size_t)b * T + t;
  *o = accumulate ? *o + acc : acc;
}

// dx (+)= coef * sign(x)
__global__ void __launch_bounds__(256)
    sign_kernel(const float *__restrict__ x, float *__restrict__ dx, long long n, float coef, int accumulate) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = x[i];
    const float g = coef * (float)((v > 0.0f) - (v < 0.0f));
    dx[i] = accumulate ? dx[i] + g : g;
  }
}

template <int MODE>
int launch_stft_bwd(const float *x0, const float *x1, int B, int T, int n_fft, int hop, float *gframes, float *dx,
                    int first_bin, float thresh, const float *fb, const int *band, int n_mels, float coef,
                    int accumulate, cudaStream_t st) {
  const int F = 1 + T / hop;
  const bool two = MODE == MODE_LOUD || MODE == MODE_MEL;
  const long long items = (long long)B * (two ? F : (F + 1) / 2);
  if (items == 0) return 0;
  if (T <= n_fft / 2) {
    set_error("stft: reflect padding needs T > n_fft/2 (T=%d, n_fft=%d)", T, n_fft);
    return -1;
  }
  if (MODE == MODE_MEL && n_mels > 128) { set_error("mel backward: n_mels must be <= 128"); return -1; }
  const int grid = (int)(items < 8LL * sm_count() ? items : 8LL * sm_count());
  switch (n_fft) {
    case 512:
      stft_bwd_kernel<512, 9, MODE><<<grid, FFT_THREADS, 0, st>>>(x0, x1, B, T, hop, F, gframes, first_bin, thresh, fb, band, n_mels, coef);
      break;
    case 1024:
      stft_bwd_kernel<1024, 10, MODE><<<grid, FFT_THREADS, 0, st>>>(x0, x1, B, T, hop, F, gframes, first_bin, thresh, fb, band, n_mels, coef);
      break;
    case 2048:
      if constexpr (MODE != MODE_MEL) {
        stft_bwd_kernel<2048, 11, MODE><<<grid, FFT_THREADS, 0, st>>>(x0, x1, B, T, hop, F, gframes, first_bin, thresh, fb, band, n_mels, coef);
      } else {   // two 2048-point buffers + the power spectra exceed the 48 KB of static shared memory
        set_error("mel backward: n_fft must be 512 or 1024");
        return -1;
      }
      break;
    default:
      set_error("stft: n_fft must be 512, 1024 or 2048 (got %d)", n_fft);
      return -1;
  }
  WM_CHECK_LAUNCH("stft_bwd");
  overlap_add_kernel<<<dim3((T + 255) / 256, B), 256, 0, st>>>(gframes, dx, T, n_fft, hop, F, accumulate);
  WM_CHECK_LAUNCH("overlap_add");
  return 0;
}

int finish(const float *partials, long long n, double scale, float *out, cudaStream_t st) {
  sum_partials_kernel<<<1, 1024, 0, st>>>(partials, n, scale, out);
  WM_CHECK_LAUNCH("sum_partials");
  return 0;
}

}  // namespace

int launch_stft_mag(const float *x, float *mag, int B, int T, int n_fft, int hop, cudaStream_t st) {
  return launch_stft<MODE_MAG>(x, nullptr, B, T, n_fft, hop, mag, nullptr, 0, 0.0f, nullptr, nullptr, 0, nullptr, st);
}

int launch_hf_penalty(const float *delta, float *out, float *partials, int B, int T, int n_fft, int first_bin,
                      cudaStream_t st) {
  int items = 0;
  const int hop = n_fft / 4, F = 1 + T / hop;
  WM_TRY(launch_stft<MODE_HF>(delta, nullptr, B, T, n_fft, hop, nullptr, partials, first_bin, 0.0f, nullptr, nullptr, 0,
                              &items, st));
  return finish(partials, items, 1.0 / ((double)B * (n_fft / 2 + 1) * F), out, st);
}

int launch_loudness(const float *clean, const float *wmk, float *out, float *partials, int B, int T, int n_fft, int hop,
                    float thresh, cudaStream_t st) {
  int items = 0;
  const int F = 1 + T / hop;
  WM_TRY(launch_stft<MODE_LOUD>(clean, wmk, B, T, n_fft, hop, nullptr, partials, 0, thresh, nullptr, nullptr, 0, &items, st));
  return finish(partials, items, 1.0 / ((double)B * (n_fft / 2 + 1) * F), out, st);
}

int launch_mel_log_l1(const float *clean, const float *wmk, const float *fb, const int *band, int n_mels, float *out,
                      float *partials, int B, int T, int n_fft, int hop, cudaStream_t st) {
  int items = 0;
  const int F = 1 + T / hop;
  WM_TRY(launch_stft<MODE_MEL>(clean, wmk, B, T, n_fft, hop, nullptr, partials, 0, 0.0f, fb, band, n_mels, &items, st));
  return finish(partials, items, 1.0 / ((double)B * n_mels * F), out, st);
}

int launch_bce_heads(const float *logits, const int64_t *message, float *loc_out, float *bce_out, float *partials,
                     int B_wm, int B2, int T, int nout, cudaStream_t st) {
  const long long rows = (long long)B2 * T;
  const int grid = (int)((rows + FFT_THREADS - 1) / FFT_THREADS < 4LL * sm_count() ? (rows + FFT_THREADS - 1) / FFT_THREADS
                                                                                   : 4LL * sm_count());
  bce_heads_kernel<<<grid, FFT_THREADS, 0, st>>>(logits, reinterpret_cast<const long long *>(message), B_wm, B2, T, nout,
                                                 partials);
  WM_CHECK_LAUNCH("bce_heads");
  WM_TRY(finish(partials, grid, 1.0 / (double)rows, loc_out, st));
  if (bce_out) {
    const double nb = (double)B_wm * T * (nout - 1);
    WM_TRY(finish(partials + grid, grid, nb > 0 ? 1.0 / nb : 0.0, bce_out, st));
  }
  return 0;
}

int launch_abs_mean(const float *x, long long n, float *out, float *partials, cudaStream_t st) {
  const int grid = (int)((n + FFT_THREADS - 1) / FFT_THREADS < 4LL * sm_count() ? (n + FFT_THREADS - 1) / FFT_THREADS
                                                                                : 4LL * sm_count());
  abs_sum_kernel<<<grid, FFT_THREADS, 0, st>>>(x, n, partials);
  WM_CHECK_LAUNCH("abs_sum");
  return finish(partials, grid, 1.0 / (double)n, out, st);
}

size_t stft_bwd_scratch_floats(int B, int T, int n_fft, int hop) { return (size_t)B * (1 + T / hop) * n_fft; }

// d_delta (+)= weight * d high_freq_penalty / d delta
int launch_hf_penalty_bwd(const float *delta, float *d_delta, float *gframes, int B, int T, int n_fft, int first_bin,
                          float weight, int accumulate, cudaStream_t st) {
  const int hop = n_fft / 4, F = 1 + T / hop;
  const float coef = (float)((double)weight / ((double)B * (n_fft / 2 + 1) * F));
  return launch_stft_bwd<MODE_HF>(delta, nullptr, B, T, n_fft, hop, gframes, d_delta, first_bin, 0.0f, nullptr, nullptr, 0,
                                  coef, accumulate, st);
}

int launch_loudness_bwd(const float *clean, const float *wmk, float *d_wmk, float *gframes, int B, int T, int n_fft,
                        int hop, float thresh, float weight, int accumulate, cudaStream_t st) {
  const int F = 1 + T / hop;
  const float coef = (float)((double)weight / ((double)B * (n_fft / 2 + 1) * F));
  return launch_stft_bwd<MODE_LOUD>(clean, wmk, B, T, n_fft, hop, gframes, d_wmk, 0, thresh, nullptr, nullptr, 0, coef,
                                    accumulate, st);
}

int launch_mel_log_l1_bwd(const float *clean, const float *wmk, const float *fb, const int *band, int n_mels,
                          float *d_wmk, float *gframes, int B, int T, int n_fft, int hop, float weight, int accumulate,
                          cudaStream_t st) {
  const int F = 1 + T / hop;
  const float coef = (float)((double)weight / ((double)B * n_mels * F));
  return launch_stft_bwd<MODE_MEL>(clean, wmk, B, T, n_fft, hop, gframes, d_wmk, 0, 0.0f, fb, band, n_mels, coef,
                                   accumulate, st);
}

int launch_abs_mean_bwd(const float *x, float *dx, long long n, float weight, int accumulate, cudaStream_t st) {
  if (n == 0) return 0;
  const int grid = (int)((n + 255) / 256 < 8LL * sm_count() ? (n + 255) / 256 : 8LL * sm_count());
  sign_kernel<<<grid, 256, 0, st>>>(x, dx, n, (float)((double)weight / (double)n), accumulate);
  WM_CHECK_LAUNCH("sign");
  return 0;
}

}  // namespace wm
