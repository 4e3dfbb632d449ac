// The data formats either side of the embed+detect path, on the device:
//   * torchaudio.transforms.Resample (py/main16.py:985,1121): windowed-sinc polyphase resampling; the kernel table is
//     built on the host exactly as torchaudio builds it, the device applies it;
//   * 16-bit PCM quantisation / de-quantisation (py/main15.py:859-860);
//   * per-file quality metrics of generate_watermarked_audio (py/main16.py:1030-1049, compute_si_snr :764-773):
//     watermark RMS, SI-SNR and power ratio from one pass of double-precision sums.
// All HBM-bound streaming kernels.
#include "wm_common.h"

namespace wm {

namespace {

// y[b][m * up + j] = sum_k kern[k][j] * x[b][m * down + k - width]   (x zero outside [0, Tin)), kern [K][up]
__global__ void __launch_bounds__(256)
    resample_kernel(const float *__restrict__ x, const float *__restrict__ kern, float *__restrict__ y, int Tin, int Tout,
                    int down, int up, int K, int width) {
  const int b = blockIdx.y;
  const float *xb = x + (size_t)b * Tin;
  for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < Tout; o += gridDim.x * blockDim.x) {
    const int m = o / up, j = o - m * up;
    const int i0 = m * down - width;
    float acc = 0.0f;
    for (int k = 0; k < K; ++k) {
      const int i = i0 + k;
      if (i >= 0 && i < Tin) acc = fmaf(__ldg(kern + (size_t)k * up + j), __ldg(xb + i), acc);
    }
    y[(size_t)b * Tout + o] = acc;
  }
}

// (clamp(x, -1, 1) * 32767) truncated toward zero, as torch's float -> int16 cast does (py/main15.py:860)
__global__ void pcm16_quantize_kernel(const float *__restrict__ x, short *__restrict__ q, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    q[i] = (short)(int)(fminf(fmaxf(x[i], -1.0f), 1.0f) * 32767.0f);
}
__global__ void pcm16_dequantize_kernel(const short *__restrict__ q, float *__restrict__ x, long long n, float scale) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    x[i] = (float)q[i] * scale;
}

// one block per row: out[b] = {watermark_rms, si_snr_db, power_ratio_db} over the first valid_len[b] samples
__global__ void __launch_bounds__(256)
    file_metrics_kernel(const float *__restrict__ s, const float *__restrict__ sw, const int *__restrict__ valid_len,
                        float *__restrict__ out, int T) {
  __shared__ double red[6][8];
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = valid_len ? min(max(valid_len[b], 0), T) : T;
  const float *a = s + (size_t)b * T, *c = sw + (size_t)b * T;
  double v[6] = {0, 0, 0, 0, 0, 0};   // sum s, sum sw, sum s^2, sum sw^2, sum s*sw, sum delta^2
  for (int i = threadIdx.x; i < n; i += 256) {
    const double p = a[i], q = c[i], d = q - p;
    v[0] += p; v[1] += q; v[2] += p * p; v[3] += q * q; v[4] += p * q; v[5] += d * d;
  }
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    for (int sft = 16; sft > 0; sft >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], sft);
    if (lane == 0) red[k][warp] = v[k];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t[6];
    for (int k = 0; k < 6; ++k) { t[k] = 0; for (int w = 0; w < 8; ++w) t[k] += red[k][w]; }
    const double N = n > 0 ? (double)n : 1.0, eps = 1e-8;
    const double ms = t[0] / N, mw = t[1] / N;
    const double norm_s = t[2] - N * ms * ms, norm_w = t[3] - N * mw * mw, dot = t[4] - N * ms * mw;
    const double alpha = dot / (norm_s + eps);
    const double tgt = alpha * alpha * norm_s, noise = norm_w - 2.0 * alpha * dot + alpha * alpha * norm_s;
    out[b * 3 + 0] = (float)sqrt(t[5] / N);
    out[b * 3 + 1] = (float)(10.0 * log10(tgt / (fmax(noise, 0.0) + eps)));
    out[b * 3 + 2] = (float)(10.0 * log10((t[2] / N) / fmax(t[5] / N, 1e-300)));
  }
}

}  // namespace

int launch_resample(const float *x, const float *kern, float *y, int B, int Tin, int Tout, int down, int up, int K,
                    int width, cudaStream_t st) {
  if (B == 0 || Tout == 0) return 0;
  dim3 grid((Tout + 255) / 256 < 4096 ? (Tout + 255) / 256 : 4096, B);
  resample_kernel<<<grid, 256, 0, st>>>(x, kern, y, Tin, Tout, down, up, K, width);
  WM_CHECK_LAUNCH("resample");
  return 0;
}

int launch_pcm16(const float *x, short *q, float *xo, long long n, int quantize, float scale, cudaStream_t st) {
  if (n == 0) return 0;
  const int grid = (int)((n + 255) / 256 < 8LL * sm_count() ? (n + 255) / 256 : 8LL * sm_count());
  if (quantize) pcm16_quantize_kernel<<<grid, 256, 0, st>>>(x, q, n);
  else pcm16_dequantize_kernel<<<grid, 256, 0, st>>>(q, xo, n, scale);
  WM_CHECK_LAUNCH("pcm16");
  return 0;
}

int launch_file_metrics(const float *s, const float *sw, const int *valid_len, float *out, int B, int T, cudaStream_t st) {
  if (B == 0) return 0;
  file_metrics_kernel<<<B, 256, 0, st>>>(s, sw, valid_len, out, T);
  WM_CHECK_LAUNCH("file_metrics");
  return 0;
}

}  // namespace wm
