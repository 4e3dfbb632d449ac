// Formats and evaluation reductions either side of the embed+detect path, on the device:
//   * the save path of py/main15.py:850-867 (and main15c.ipynb cell 4): torchaudio.functional.lowpass_biquad at 7 kHz
//     (a second-order IIR, lfilter semantics incl. its clamp to [-1, 1]) followed by 16-bit PCM quantisation;
//   * detection statistics of the evaluation cells: confusion counts at a threshold (py/main16.py:1335-1341,
//     prediction = p >= thresh), ROC operating points and the area under the ROC curve (py/main16.py:2372-2386).
//
// The biquad is a linear recurrence along time.  A file is cut into chunks of BQ_L samples; every chunk is filtered
// from a zero state by one thread (y_zs), a single thread per file then carries the 2-vector state across chunks
// (state' = P state + chunk's final zero-state pair, P = the chunk's transition matrix), and a third pass adds each
// chunk's zero-input response h1[n] s1 + h2[n] s2.  All recurrences run in double precision on the fp32 coefficient
// values torchaudio would use, so the result differs from torchaudio's sequential fp32 loop only by that loop's own
// rounding (tests: <= 2e-6; PCM codes equal up to 1 LSB where the two values straddle an integer, ~1 % of samples).
#include "wm_common.h"

namespace wm {

namespace {

constexpr int BQ_L = 64;          // samples per chunk
constexpr int BQ_TPB = 128;       // chunks (threads) per block; shared tile 128 x (64 + 1) floats

struct BiquadCoef {
  double b0, b1, b2, a1, a2;      // already divided by a0
};

// pass 1: zero-state response of every chunk; final pair (y[L-1], y[L-2]) of the chunk -> states
__global__ void __launch_bounds__(BQ_TPB)
    biquad_chunk_kernel(const float *__restrict__ x, float *__restrict__ y, double2 *__restrict__ zs_final, long long N,
                        long long nchunk, BiquadCoef k) {
  __shared__ float tile[BQ_TPB][BQ_L + 1];
  const long long row = blockIdx.y;
  const float *xr = x + row * N;
  float *yr = y + row * N;
  const long long c0 = (long long)blockIdx.x * BQ_TPB;
  const long long base = c0 * BQ_L;
  // coalesced load of the block's 128 x 64 samples (+ the two samples in front of every chunk come from the tile / global)
  for (int i = threadIdx.x; i < BQ_TPB * BQ_L; i += BQ_TPB) {
    const long long g = base + i;
    tile[i / BQ_L][i % BQ_L] = g < N ? xr[g] : 0.0f;
  }
  __syncthreads();
  const long long c = c0 + threadIdx.x;
  const long long n0 = c * BQ_L;
  // the two input samples in front of the chunk: the left neighbour's row of the tile (read before anybody
  // overwrites a row with outputs) or, for the block's first chunk, global memory
  double xm1 = 0.0, xm2 = 0.0;
  if (c < nchunk) {
    if (n0 >= 1) xm1 = (double)(threadIdx.x > 0 ? tile[threadIdx.x - 1][BQ_L - 1] : xr[n0 - 1]);
    if (n0 >= 2) xm2 = (double)(threadIdx.x > 0 ? tile[threadIdx.x - 1][BQ_L - 2] : xr[n0 - 2]);
  }
  __syncthreads();
  if (c < nchunk) {
    double y1 = 0.0, y2 = 0.0;
#pragma unroll 4
    for (int i = 0; i < BQ_L; ++i) {
      const double xi = tile[threadIdx.x][i];
      const double yi = k.b0 * xi + k.b1 * xm1 + k.b2 * xm2 - k.a1 * y1 - k.a2 * y2;
      xm2 = xm1; xm1 = xi;
      y2 = y1; y1 = yi;
      tile[threadIdx.x][i] = (float)yi;
    }
    zs_final[row * nchunk + c] = make_double2(y1, y2);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < BQ_TPB * BQ_L; i += BQ_TPB) {
    const long long g = base + i;
    if (g < N) yr[g] = tile[i / BQ_L][i % BQ_L];
  }
}

// pass 2: one thread per file walks the chunks: state entering chunk c+1 = P * (state entering c) + zs_final[c]
__global__ void biquad_carry_kernel(double2 *__restrict__ zs_final, long long nchunk, int rows, double p11, double p12,
                                    double p21, double p22) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= rows) return;
  double2 *st = zs_final + (long long)row * nchunk;
  double s1 = 0.0, s2 = 0.0;                 // (y[-1], y[-2]) entering the chunk
  for (long long c = 0; c < nchunk; ++c) {
    const double2 z = st[c];
    st[c] = make_double2(s1, s2);            // overwrite with the state ENTERING chunk c
    const double n1 = p11 * s1 + p12 * s2 + z.x, n2 = p21 * s1 + p22 * s2 + z.y;
    s1 = n1; s2 = n2;
  }
}

// pass 3: y += h1[i] s1 + h2[i] s2 (zero-input response of the entering state), clamp as lfilter does, optional PCM16
__global__ void __launch_bounds__(256)
    biquad_finish_kernel(float *__restrict__ y, short *__restrict__ q, const double2 *__restrict__ state_in,
                         const double *__restrict__ h, long long N, long long nchunk, int clamp) {
  const long long row = blockIdx.y;
  float *yr = y + row * N;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < N; g += (long long)gridDim.x * blockDim.x) {
    const long long c = g / BQ_L;
    const int i = (int)(g - c * BQ_L);
    const double2 s = state_in[row * nchunk + c];
    float v = (float)((double)yr[g] + h[i] * s.x + h[BQ_L + i] * s.y);
    if (clamp) v = fminf(fmaxf(v, -1.0f), 1.0f);
    yr[g] = v;
    if (q != nullptr) q[row * N + g] = (short)(int)(fminf(fmaxf(v, -1.0f), 1.0f) * 32767.0f);
  }
}

// out[0..3] = tn, fp, fn, tp with prediction = score >= thresh (py/main16.py:1337)
__global__ void __launch_bounds__(256)
    confusion_kernel(const float *__restrict__ clean, long long n0, const float *__restrict__ wm, long long n1, float thresh,
                     unsigned long long *__restrict__ out) {
  unsigned long long fp = 0, tp = 0;
  const long long stride = (long long)gridDim.x * blockDim.x, first = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (long long i = first; i < n0; i += stride) fp += clean[i] >= thresh;
  for (long long i = first; i < n1; i += stride) tp += wm[i] >= thresh;
  for (int sft = 16; sft > 0; sft >>= 1) {
    fp += __shfl_xor_sync(0xffffffffu, fp, sft);
    tp += __shfl_xor_sync(0xffffffffu, tp, sft);
  }
  if ((threadIdx.x & 31) == 0) {      // integer sums: the order of the atomics does not change the result
    atomicAdd(out + 1, fp);
    atomicAdd(out + 3, tp);
  }
  if (first == 0) { atomicAdd(out + 0, (unsigned long long)n0); atomicAdd(out + 2, (unsigned long long)n1); }
}
__global__ void confusion_finish_kernel(unsigned long long *out) {   // tn = n0 - fp, fn = n1 - tp
  out[0] -= out[1];
  out[2] -= out[3];
}

// ROC operating points: for every threshold t, fp[t] = #{clean >= t}, tp[t] = #{wm >= t}; one block per threshold
__global__ void __launch_bounds__(256)
    roc_points_kernel(const float *__restrict__ clean, long long n0, const float *__restrict__ wm, long long n1,
                      const float *__restrict__ thr, int *__restrict__ fp, int *__restrict__ tp) {
  __shared__ int red[2][8];
  const float t = thr[blockIdx.x];
  int f = 0, p = 0;
  for (long long i = threadIdx.x; i < n0; i += 256) f += clean[i] >= t;
  for (long long i = threadIdx.x; i < n1; i += 256) p += wm[i] >= t;
  for (int sft = 16; sft > 0; sft >>= 1) {
    f += __shfl_xor_sync(0xffffffffu, f, sft);
    p += __shfl_xor_sync(0xffffffffu, p, sft);
  }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = f; red[1][threadIdx.x >> 5] = p; }
  __syncthreads();
  if (threadIdx.x == 0) {
    int a = 0, b = 0;
    for (int w = 0; w < 8; ++w) { a += red[0][w]; b += red[1][w]; }
    fp[blockIdx.x] = a;
    tp[blockIdx.x] = b;
  }
}

// area under the ROC curve as the rank statistic: (#{wm > clean} + #{wm == clean} / 2) / (n0 n1), which is what the
// trapezoid rule gives on the full curve; out[0] += 2 * wins + ties (integer), one block row per slab of wm scores
__global__ void __launch_bounds__(256)
    auc_pairs_kernel(const float *__restrict__ clean, long long n0, const float *__restrict__ wm, long long n1,
                     unsigned long long *__restrict__ out) {
  unsigned long long acc = 0;
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n1; j += (long long)gridDim.x * blockDim.x) {
    const float w = wm[j];
    for (long long i = 0; i < n0; ++i) {
      const float c = __ldg(clean + i);
      acc += w > c ? 2u : (w == c ? 1u : 0u);
    }
  }
  for (int sft = 16; sft > 0; sft >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, sft);
  if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}

}  // namespace

size_t biquad_scratch_bytes(int rows, long long N) {
  const long long nchunk = (N + BQ_L - 1) / BQ_L;
  return (size_t)rows * nchunk * sizeof(double2) + 2 * BQ_L * sizeof(double);
}

// x, y: [rows][N] fp32 (NOT in place: a block's first chunk looks two input samples back into its neighbour's range); q: optional [rows][N] int16; b / a: the filter as lfilter takes it (a[0] != 0)
int launch_biquad(const float *x, float *y, short *q, int rows, long long N, const double *b, const double *a, int clamp,
                  void *scratch, cudaStream_t st) {
  if (rows == 0 || N == 0) return 0;
  const long long nchunk = (N + BQ_L - 1) / BQ_L;
  BiquadCoef k{b[0] / a[0], b[1] / a[0], b[2] / a[0], a[1] / a[0], a[2] / a[0]};
  // chunk transition matrix P = M^L, M = [[-a1, -a2], [1, 0]] acting on (y[n-1], y[n-2]); zero-input responses
  // h1 (state (1,0)) and h2 (state (0,1)), L samples each -- a few hundred flops on the host
  double h[2 * BQ_L], P[4];
  for (int col = 0; col < 2; ++col) {
    double y1 = col == 0 ? 1.0 : 0.0, y2 = col == 0 ? 0.0 : 1.0;
    for (int i = 0; i < BQ_L; ++i) {
      const double yi = -k.a1 * y1 - k.a2 * y2;
      y2 = y1; y1 = yi;
      h[col * BQ_L + i] = yi;
    }
    P[0 + col] = y1;   // row 0: new y[-1]
    P[2 + col] = y2;   // row 1: new y[-2]
  }
  double2 *states = reinterpret_cast<double2 *>(scratch);
  double *h_dev = reinterpret_cast<double *>(states + (size_t)rows * nchunk);
  WM_CHECK_CUDA(cudaMemcpyAsync(h_dev, h, sizeof(h), cudaMemcpyHostToDevice, st));
  dim3 g1((unsigned)((nchunk + BQ_TPB - 1) / BQ_TPB), rows);
  biquad_chunk_kernel<<<g1, BQ_TPB, 0, st>>>(x, y, states, N, nchunk, k);
  WM_CHECK_LAUNCH("biquad_chunk");
  biquad_carry_kernel<<<(rows + 63) / 64, 64, 0, st>>>(states, nchunk, rows, P[0], P[1], P[2], P[3]);
  WM_CHECK_LAUNCH("biquad_carry");
  long long blocks = (N + 255) / 256;
  dim3 g3((unsigned)(blocks < 4 * sm_count() ? blocks : 4 * sm_count()), rows);
  biquad_finish_kernel<<<g3, 256, 0, st>>>(y, q, states, h_dev, N, nchunk, clamp);
  WM_CHECK_LAUNCH("biquad_finish");
  return 0;
}

int launch_confusion(const float *clean, long long n0, const float *wm, long long n1, float thresh, unsigned long long *out4,
                     cudaStream_t st) {
  WM_CHECK_CUDA(cudaMemsetAsync(out4, 0, 4 * sizeof(unsigned long long), st));
  long long n = n0 > n1 ? n0 : n1;
  int blocks = (int)((n + 255) / 256 < 2 * sm_count() ? (n + 255) / 256 : 2 * sm_count());
  if (blocks < 1) blocks = 1;
  confusion_kernel<<<blocks, 256, 0, st>>>(clean, n0, wm, n1, thresh, out4);
  WM_CHECK_LAUNCH("confusion");
  confusion_finish_kernel<<<1, 1, 0, st>>>(out4);
  WM_CHECK_LAUNCH("confusion_finish");
  return 0;
}

int launch_roc_points(const float *clean, long long n0, const float *wm, long long n1, const float *thr, int nt, int *fp,
                      int *tp, cudaStream_t st) {
  if (nt == 0) return 0;
  roc_points_kernel<<<nt, 256, 0, st>>>(clean, n0, wm, n1, thr, fp, tp);
  WM_CHECK_LAUNCH("roc_points");
  return 0;
}

int launch_auc_pairs(const float *clean, long long n0, const float *wm, long long n1, unsigned long long *out, cudaStream_t st) {
  WM_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(unsigned long long), st));
  if (n0 == 0 || n1 == 0) return 0;
  int blocks = (int)((n1 + 255) / 256 < 8 * sm_count() ? (n1 + 255) / 256 : 8 * sm_count());
  auc_pairs_kernel<<<blocks, 256, 0, st>>>(clean, n0, wm, n1, out);
  WM_CHECK_LAUNCH("auc_pairs");
  return 0;
}

}  // namespace wm
