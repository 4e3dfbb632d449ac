// fp32 CUDA-core convolution kernels of the main16 path (sm_100a).
//   conv_in_k7   Conv1d(1,64,7,p=3)            py/main16.py:134,177
//   conv64_fp32  Conv1d(64,64,3,p=1)+BN(+ReLU) py/main16.py:116-121  (taps = 3)
//                ConvTranspose1d(64,64,7,p=3)  py/main16.py:144      (taps = 7, pre-flipped)
//   head         Conv1d(64,nout,1)             py/main16.py:146,180
// The fp32 conv64 kernel is the exact-arithmetic path (WM_MATH_FP32) and the on-GPU
// cross-check of the tcgen05 kernel in wm_conv_tc.cu.
#include "wm_common.h"

namespace wm {

// ---------------------------------------------------------------------------
// Conv1d(1,64,7,p=3): y[b][t][c] = bias[c] + sum_j w[j][c] * s[b][t+j-3]
// One thread produces 4 channels of one time step (float4 store, 256 B per row
// coalesced over 16 lanes).  HBM-bound: 4 B read, 256 B written per time step.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) conv_in_k7_kernel(const float *__restrict__ s,
                                                          const float *__restrict__ w,
                                                          const float *__restrict__ bias,
                                                          float *__restrict__ y, int T) {
  constexpr int TT = 128;  // time steps per block
  __shared__ __align__(16) float ss[TT + 8];
  __shared__ __align__(16) float ws[7 * 64 + 64];
  const int b = blockIdx.y, t0 = blockIdx.x * TT, tid = threadIdx.x;
  const float *sb = s + (size_t)b * T;
  for (int i = tid; i < TT + 6; i += 256) {
    int t = t0 + i - 3;
    ss[i] = (t >= 0 && t < T) ? sb[t] : 0.0f;
  }
  for (int i = tid; i < 7 * 64; i += 256) ws[i] = w[i];
  if (tid < 64) ws[7 * 64 + tid] = bias[tid];
  __syncthreads();
  const int c0 = (tid & 15) * 4, r0 = tid >> 4;
  float4 wv[7];
#pragma unroll
  for (int j = 0; j < 7; ++j) wv[j] = *reinterpret_cast<const float4 *>(&ws[j * 64 + c0]);
  const float4 bv = *reinterpret_cast<const float4 *>(&ws[7 * 64 + c0]);
#pragma unroll
  for (int i = 0; i < TT / 16; ++i) {
    int r = r0 + 16 * i, t = t0 + r;
    if (t >= T) break;
    float4 a = bv;
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      float v = ss[r + j];
      a.x = fmaf(v, wv[j].x, a.x);
      a.y = fmaf(v, wv[j].y, a.y);
      a.z = fmaf(v, wv[j].z, a.z);
      a.w = fmaf(v, wv[j].w, a.w);
    }
    *reinterpret_cast<float4 *>(&y[((size_t)b * T + t) * 64 + c0]) = a;
  }
}

int launch_conv_in_k7(const float *s, const float *w, const float *b, float *y, int B, int T,
                      cudaStream_t st) {
  if (B == 0 || T == 0) return 0;
  dim3 grid((T + 127) / 128, B);
  conv_in_k7_kernel<<<grid, 256, 0, st>>>(s, w, b, y, T);
  WM_CHECK_LAUNCH("conv_in_k7");
  return 0;
}

// ---------------------------------------------------------------------------
// 64 -> 64 convolution, fp32 FMA.  Block = 128 time steps x 64 output channels,
// 256 threads, each thread 8 time steps (strided by 16) x 4 channels.
// smem: x tile (128 + TAPS - 1) rows padded to 68 floats, one tap of weights
// (64 x 64) at a time.
// ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long pack2f(float a, float b) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void unpack2f(unsigned long long v, float &a, float &b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ unsigned long long fma2f(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

template <int TAPS>
__global__ void __launch_bounds__(256)
    conv64_fp32_kernel(const float *__restrict__ x, const float *__restrict__ w,
                       const float *__restrict__ bias, const float *__restrict__ residual,
                       const float *__restrict__ chan_add, float *__restrict__ y, int T, int relu) {
  constexpr int TT = 128, P = TAPS / 2, ROWS = TT + TAPS - 1, LD = 68;
  extern __shared__ __align__(16) float smem[];
  float *xs = smem;              // [ROWS][LD]
  float *ws = smem + ROWS * LD;  // [64][64]
  const int b = blockIdx.y, t0 = blockIdx.x * TT, tid = threadIdx.x;
  const float *xb = x + (size_t)b * T * 64;

  // x tile (zero outside [0,T); the per-clip channel vector is added to in-range rows only)
  for (int i = tid; i < ROWS * 16; i += 256) {
    int r = i >> 4, c4 = (i & 15) * 4, t = t0 + r - P;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t >= 0 && t < T) {
      v = *reinterpret_cast<const float4 *>(&xb[(size_t)t * 64 + c4]);
      if (chan_add) {
        float4 e = *reinterpret_cast<const float4 *>(&chan_add[(size_t)b * 64 + c4]);
        v.x += e.x; v.y += e.y; v.z += e.z; v.w += e.w;
      }
    }
    *reinterpret_cast<float4 *>(&xs[r * LD + c4]) = v;
  }

  const int tx = tid & 15, ty = tid >> 4, co0 = tx * 4;
  // accumulators as packed fp32 pairs (fma.rn.f32x2: two FMAs per issued instruction): [time step][channel pair]
  unsigned long long acc2[8][2];
  {
    float4 bv = *reinterpret_cast<const float4 *>(&bias[co0]);
    const unsigned long long b01 = pack2f(bv.x, bv.y), b23 = pack2f(bv.z, bv.w);
#pragma unroll
    for (int i = 0; i < 8; ++i) { acc2[i][0] = b01; acc2[i][1] = b23; }
  }

  for (int j = 0; j < TAPS; ++j) {
    __syncthreads();  // previous tap's weights consumed (and x tile visible on j == 0)
    for (int i = tid; i < 64 * 16; i += 256)
      *reinterpret_cast<float4 *>(&ws[i * 4]) = *reinterpret_cast<const float4 *>(&w[(size_t)j * 4096 + i * 4]);
    __syncthreads();
#pragma unroll 4
    for (int c4 = 0; c4 < 16; ++c4) {
      ulonglong2 wv[4];      // weights of input channels c4*4 + q for this thread's 4 output channels, as two pairs
#pragma unroll
      for (int q = 0; q < 4; ++q) wv[q] = *reinterpret_cast<const ulonglong2 *>(&ws[(c4 * 4 + q) * 64 + co0]);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float4 xv = *reinterpret_cast<const float4 *>(&xs[(ty + 16 * i + j) * LD + c4 * 4]);
        const unsigned long long x0 = pack2f(xv.x, xv.x), x1 = pack2f(xv.y, xv.y), x2 = pack2f(xv.z, xv.z),
                                 x3 = pack2f(xv.w, xv.w);
        acc2[i][0] = fma2f(x0, wv[0].x, acc2[i][0]); acc2[i][1] = fma2f(x0, wv[0].y, acc2[i][1]);
        acc2[i][0] = fma2f(x1, wv[1].x, acc2[i][0]); acc2[i][1] = fma2f(x1, wv[1].y, acc2[i][1]);
        acc2[i][0] = fma2f(x2, wv[2].x, acc2[i][0]); acc2[i][1] = fma2f(x2, wv[2].y, acc2[i][1]);
        acc2[i][0] = fma2f(x3, wv[3].x, acc2[i][0]); acc2[i][1] = fma2f(x3, wv[3].y, acc2[i][1]);
      }
    }
  }

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    unpack2f(acc2[i][0], acc[i][0], acc[i][1]);
    unpack2f(acc2[i][1], acc[i][2], acc[i][3]);
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int t = t0 + ty + 16 * i;
    if (t >= T) continue;
    size_t o = ((size_t)b * T + t) * 64 + co0;
    float4 v = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    if (residual) {
      float4 r = *reinterpret_cast<const float4 *>(&residual[o]);
      v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
    }
    if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    *reinterpret_cast<float4 *>(&y[o]) = v;
  }
}

template <int TAPS>
static int launch_conv64_t(const float *x, const float *w, const float *bias, const float *residual,
                           const float *chan_add, float *y, int B, int T, int relu, cudaStream_t st) {
  constexpr int SMEM = ((128 + TAPS - 1) * 68 + 64 * 64) * 4;
  static bool attr_set = false;
  if (!attr_set) {
    WM_CHECK_CUDA(cudaFuncSetAttribute(conv64_fp32_kernel<TAPS>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    attr_set = true;
  }
  dim3 grid((T + 127) / 128, B);
  conv64_fp32_kernel<TAPS><<<grid, 256, SMEM, st>>>(x, w, bias, residual, chan_add, y, T, relu);
  WM_CHECK_LAUNCH("conv64_fp32");
  return 0;
}

int launch_conv64_fp32(const float *x, const float *w, const float *bias, const float *residual,
                       const float *chan_add, float *y, int B, int T, int taps, int relu,
                       cudaStream_t st) {
  if (B == 0 || T == 0) return 0;
  switch (taps) {
    case 1: return launch_conv64_t<1>(x, w, bias, residual, chan_add, y, B, T, relu, st);
    case 3: return launch_conv64_t<3>(x, w, bias, residual, chan_add, y, B, T, relu, st);
    case 7: return launch_conv64_t<7>(x, w, bias, residual, chan_add, y, B, T, relu, st);
    default: set_error("conv64: taps must be 1, 3 or 7 (got %d)", taps); return -1;
  }
}

// ---------------------------------------------------------------------------
// Conv1d(64,nout,1): y[b][t][o] = bias[o] + sum_c w[o][c] x[b][t][c].
// A warp handles 32 consecutive rows, one row per lane.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) head_kernel(const float *__restrict__ x,
                                                    const float *__restrict__ w,
                                                    const float *__restrict__ bias,
                                                    float *__restrict__ y, long long rows, int nout) {
  __shared__ __align__(16) float ws[WM_MAX_HEAD * 64];
  __shared__ float bs[WM_MAX_HEAD];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < nout * 64; i += 256) ws[i] = w[i];
  if (tid < nout) bs[tid] = bias[tid];
  __syncthreads();
  for (long long r0 = ((long long)blockIdx.x * 8 + warp) * 32; r0 < rows;
       r0 += (long long)gridDim.x * 256) {
    // each lane owns one 256-byte row (two full cache lines, consumed by this lane alone)
    long long rr = r0 + lane < rows ? r0 + lane : rows - 1;
    float xr[64];
#pragma unroll
    for (int c4 = 0; c4 < 16; ++c4) {
      float4 v = __ldg(reinterpret_cast<const float4 *>(&x[rr * 64 + c4 * 4]));
      xr[c4 * 4] = v.x; xr[c4 * 4 + 1] = v.y; xr[c4 * 4 + 2] = v.z; xr[c4 * 4 + 3] = v.w;
    }
    if (r0 + lane < rows) {
      for (int o = 0; o < nout; ++o) {
        float a = bs[o];
#pragma unroll
        for (int c = 0; c < 64; ++c) a = fmaf(xr[c], ws[o * 64 + c], a);
        y[(r0 + lane) * nout + o] = a;
      }
    }
  }
}

int launch_head(const float *x, const float *w, const float *b, float *y, int B, int T, int nout,
                cudaStream_t st) {
  if (B == 0 || T == 0) return 0;
  if (nout < 1 || nout > WM_MAX_HEAD) { set_error("head: nout must be in [1,%d]", WM_MAX_HEAD); return -1; }
  long long rows = (long long)B * T;
  long long blocks = (rows + 255) / 256;
  int grid = (int)(blocks < (long long)sm_count() * 8 ? blocks : (long long)sm_count() * 8);
  head_kernel<<<grid, 256, 0, st>>>(x, w, b, y, rows, nout);
  WM_CHECK_LAUNCH("head");
  return 0;
}

// embedding rows: out[b][0..63] = table[idx[b]][0..63]   (py/main16.py:156-158)
__global__ void gather_rows_kernel(const float *__restrict__ table, long long rows,
                                   const int64_t *__restrict__ idx, float *__restrict__ out, int B) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * 64) return;
  int b = i >> 6, c = i & 63;
  long long r = idx[b];
  out[i] = (r >= 0 && r < rows) ? table[r * 64 + c] : 0.0f;
}

int launch_gather_rows(const float *table, int64_t rows, const int64_t *idx, float *out, int B,
                       cudaStream_t st) {
  if (B == 0) return 0;
  gather_rows_kernel<<<(B * 64 + 255) / 256, 256, 0, st>>>(table, rows, idx, out, B);
  WM_CHECK_LAUNCH("gather_rows");
  return 0;
}

}  // namespace wm
