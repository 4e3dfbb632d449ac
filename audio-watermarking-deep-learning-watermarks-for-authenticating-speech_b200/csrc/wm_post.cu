// HBM-bound tail kernels of the main16 path (sm_100a):
//   postprocess   fir_lowpass -> clamp_peak -> limit_rms -> s + delta   py/main16.py:53-72,245-248
//   detect_heads  sigmoid(ch 0), per-clip mean, mean message logits, votes  py/main16.py:1142-1146,393-398
//   head_detect   the same straight from the last 64-channel activation (1x1 head fused)
#include "wm_common.h"

namespace wm {

namespace {
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// deterministic block sum (fixed tree); result valid in every thread
template <int NT>
__device__ __forceinline__ float block_sum(float v, float *red /*[NT/32]*/) {
  v = warp_sum(v);
  __syncthreads();  // protect `red` from a previous use
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = 0.0f;
#pragma unroll
  for (int i = 0; i < NT / 32; ++i) s += red[i];
  return s;
}
}  // namespace

// ---------------------------------------------------------------------------
// One block per clip.  mode != 0 keeps the raw delta (with a 50-sample zero halo) and
// the filtered+clamped delta in shared memory, so the clip is read once and the
// outputs written once: 8 B read + 8 B written per sample.
// ---------------------------------------------------------------------------
constexpr int kPostThreads = 1024;

__global__ void __launch_bounds__(kPostThreads)
    postprocess_kernel(const float *__restrict__ delta_raw, const float *__restrict__ s,
                       const float *__restrict__ fir, float *__restrict__ delta,
                       float *__restrict__ s_w, float *__restrict__ rms_out, int T, int mode,
                       float peak, float max_rms, float eps) {
  extern __shared__ __align__(16) float psm[];
  __shared__ float red[kPostThreads / 32];
  __shared__ __align__(16) float taps[(WM_FIR_TAPS + 3) / 4 * 4];
  constexpr int H = WM_FIR_TAPS / 2;
  const int b = blockIdx.x, tid = threadIdx.x;
  const float *dr = delta_raw + (size_t)b * T;
  float sumsq = 0.0f;

  if (mode != 0) {
    const bool do_fir = mode & 1, do_clamp = mode & 2, do_rms = mode & 4;
    constexpr int KP = (WM_FIR_TAPS + 3) / 4 * 4;      // taps padded with zeros to a multiple of 4
    constexpr int RPAD = 24;                           // zero samples behind the halo: every windowed read is in range
    float *raw = psm;                                  // [T + 2H + RPAD]
    float *flt = psm + (T + 2 * H + RPAD + 3) / 4 * 4;  // [T]
    if (tid < KP) taps[tid] = tid < WM_FIR_TAPS ? (do_fir ? fir[tid] : (tid == H ? 1.0f : 0.0f)) : 0.0f;
    for (int i = tid; i < T + 2 * H + RPAD; i += kPostThreads) {
      int t = i - H;
      raw[i] = (t >= 0 && t < T) ? dr[t] : 0.0f;
    }
    __syncthreads();
    // 8 consecutive outputs per thread: out[t] = sum_k taps[k] * raw[t + k] (raw is offset by H), taps in increasing
    // order for every output.  The 11-sample window of four taps lives in registers and advances by one 16-byte
    // shared-memory load per four taps (one scalar, bank-conflicted load per tap made this kernel LSU-bound:
    // 1.09 ms for 4096 clips at 8 % of the HBM rate).
    for (int t0 = tid * 8; t0 < T; t0 += kPostThreads * 8) {
      float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (do_fir) {
        float w[12];
        {
          const float4 lo = *reinterpret_cast<const float4 *>(&raw[t0]), hi = *reinterpret_cast<const float4 *>(&raw[t0 + 4]);
          w[0] = lo.x; w[1] = lo.y; w[2] = lo.z; w[3] = lo.w; w[4] = hi.x; w[5] = hi.y; w[6] = hi.z; w[7] = hi.w;
        }
#pragma unroll 2
        for (int kb = 0; kb < KP; kb += 4) {
          const float4 nx = *reinterpret_cast<const float4 *>(&raw[t0 + kb + 8]);
          const float4 tp = *reinterpret_cast<const float4 *>(&taps[kb]);
          w[8] = nx.x; w[9] = nx.y; w[10] = nx.z; w[11] = nx.w;
          const float tk[4] = {tp.x, tp.y, tp.z, tp.w};
#pragma unroll
          for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = fmaf(tk[j], w[i + j], a[i]);
#pragma unroll
          for (int i = 0; i < 8; ++i) w[i] = w[i + 4];
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = raw[t0 + H + i];
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        if (t0 + q < T) {
          float d = do_clamp ? fminf(fmaxf(a[q], -peak), peak) : a[q];
          flt[t0 + q] = d;
          sumsq = fmaf(d, d, sumsq);
        }
      }
    }
    float tot = block_sum<kPostThreads>(sumsq, red);
    float ms = tot / (float)T;
    float gain = do_rms ? fminf(max_rms / sqrtf(ms + eps), 1.0f) : 1.0f;
    if (tid == 0 && rms_out) rms_out[b] = gain * sqrtf(ms);
    const float *sb = s + (size_t)b * T;
    for (int t = tid; t < T; t += kPostThreads) {
      float d = flt[t] * gain;
      if (delta) delta[(size_t)b * T + t] = d;
      if (s_w) s_w[(size_t)b * T + t] = sb[t] + d;   // s is only read when s_w is wanted
    }
  } else {
    const float *sb = s + (size_t)b * T;
    for (int t = tid; t < T; t += kPostThreads) {
      float d = dr[t];
      sumsq = fmaf(d, d, sumsq);
      if (delta) delta[(size_t)b * T + t] = d;
      if (s_w) s_w[(size_t)b * T + t] = sb[t] + d;
    }
    if (rms_out) {
      float tot = block_sum<kPostThreads>(sumsq, red);
      if (tid == 0) rms_out[b] = sqrtf(tot / (float)T);
    }
  }
}

int launch_postprocess(const float *delta_raw, const float *s, const float *fir, float *delta,
                       float *s_w, float *rms_out, int B, int T, int mode, float peak,
                       float max_rms, float eps, cudaStream_t st) {
  if (B == 0 || T == 0) return 0;
  size_t smem = 0;
  if (mode < 0 || mode > 7) { set_error("postprocess: mode must be a bit mask in [0,7]"); return -1; }
  if (mode != 0) {
    if ((mode & 1) && !fir) { set_error("postprocess: fir taps required when bit 0 of mode is set"); return -1; }
    smem = (size_t)(2 * T + WM_FIR_TAPS + 32) * sizeof(float);
    if (smem > 220 * 1024) { set_error("postprocess: T=%d too long for the one-block-per-clip kernel", T); return -1; }
    static size_t attr = 0;
    if (smem > attr) {
      WM_CHECK_CUDA(cudaFuncSetAttribute(postprocess_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr = smem;
    }
  }
  postprocess_kernel<<<B, kPostThreads, smem, st>>>(delta_raw, s, fir, delta, s_w, rms_out, T, mode,
                                                    peak, max_rms, eps);
  WM_CHECK_LAUNCH("postprocess");
  return 0;
}

// ---------------------------------------------------------------------------
// Backward of clamp_peak -> limit_rms (py/main16.py:66-72) on the filtered delta d1 (= fir_lowpass(delta_raw)):
//   d2 = clamp(d1), rms = sqrt(mean(d2^2) + eps), gain = min(max_rms / rms, 1), delta = gain * d2
//   dL/dd2 = gain * g - [gain < 1] * d2 * gain / rms^2 * mean(g * d2);   dL/dd1 = dL/dd2 * [|d1| <= peak]
// One block per clip.  The FIR's own backward is the same FIR (its taps are symmetric) applied to dL/dd1.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kPostThreads)
    postprocess_bwd_kernel(const float *__restrict__ g, const float *__restrict__ d1, float *__restrict__ dd1, int T,
                           int mode, float peak, float max_rms, float eps) {
  __shared__ float red[kPostThreads / 32];
  const int b = blockIdx.x, tid = threadIdx.x;
  const bool do_clamp = mode & 2, do_rms = mode & 4;
  const float *gb = g + (size_t)b * T, *db = d1 + (size_t)b * T;
  float sumsq = 0.0f, dot = 0.0f;
  for (int t = tid; t < T; t += kPostThreads) {
    const float v = db[t];
    const float d2 = do_clamp ? fminf(fmaxf(v, -peak), peak) : v;
    sumsq = fmaf(d2, d2, sumsq);
    dot = fmaf(gb[t], d2, dot);
  }
  const float ms = block_sum<kPostThreads>(sumsq, red) / (float)T;
  const float gd = block_sum<kPostThreads>(dot, red) / (float)T;
  const float rms2 = ms + eps, ratio = max_rms / sqrtf(rms2);
  const bool scaled = do_rms && ratio < 1.0f;
  const float gain = scaled ? ratio : 1.0f;
  const float back = scaled ? gain / rms2 * gd : 0.0f;
  for (int t = tid; t < T; t += kPostThreads) {
    const float v = db[t];
    const bool inside = !do_clamp || (v >= -peak && v <= peak);
    const float d2 = do_clamp ? fminf(fmaxf(v, -peak), peak) : v;
    dd1[(size_t)b * T + t] = inside ? gain * gb[t] - back * d2 : 0.0f;
  }
}

// g = dL/d delta, d1 = the filtered (pre-clamp) delta -> d_delta_raw; scratch: B*T floats
int launch_postprocess_bwd(const float *g, const float *d1, const float *fir, float *d_delta_raw, float *scratch, int B,
                           int T, int mode, float peak, float max_rms, float eps, cudaStream_t st) {
  if (B == 0 || T == 0) return 0;
  if (mode < 0 || mode > 7) { set_error("postprocess_bwd: mode must be a bit mask in [0,7]"); return -1; }
  float *dd1 = (mode & 1) ? scratch : d_delta_raw;
  postprocess_bwd_kernel<<<B, kPostThreads, 0, st>>>(g, d1, dd1, T, mode, peak, max_rms, eps);
  WM_CHECK_LAUNCH("postprocess_bwd");
  if (mode & 1) return launch_postprocess(dd1, nullptr, fir, d_delta_raw, nullptr, nullptr, B, T, WM_POST_FIR, peak, max_rms, eps, st);
  return 0;
}

// ---------------------------------------------------------------------------
// Heads.  One block (256 threads) per clip; every thread walks rows tid, tid+256, ...
// and keeps its partial sums in registers; fixed-order block reduction at the end.
// FROM_X: logits are computed on the fly from x[b][t][64] with the 1x1 head weights.
// ---------------------------------------------------------------------------
template <bool FROM_X>
__global__ void __launch_bounds__(256)
    detect_heads_kernel(const float *__restrict__ in, const float *__restrict__ w,
                        const float *__restrict__ bias, const int *__restrict__ valid_len,
                        float *__restrict__ probs, float *__restrict__ clip_prob,
                        float *__restrict__ msg_logits, float *__restrict__ vote_frac, int T, int nout) {
  __shared__ __align__(16) float ws[FROM_X ? WM_MAX_HEAD * 64 : 4];
  __shared__ float bs[WM_MAX_HEAD];
  __shared__ float red[8];
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nbits = nout - 1;
  const int valid = valid_len ? min(max(valid_len[b], 0), T) : T;
  if (FROM_X) {
    for (int i = tid; i < nout * 64; i += 256) ws[i] = w[i];
    if (tid < nout) bs[tid] = bias[tid];
    __syncthreads();
  }
  float sp = 0.0f, sl[WM_MAX_HEAD - 1], sv[WM_MAX_HEAD - 1];
#pragma unroll
  for (int o = 0; o < WM_MAX_HEAD - 1; ++o) { sl[o] = 0.0f; sv[o] = 0.0f; }

  for (int r0 = warp * 32; r0 < T; r0 += 256) {
    const int t = r0 + lane;
    float lg[WM_MAX_HEAD];
    if (FROM_X) {
      // each lane owns one 256-byte row (two full cache lines, consumed by this lane alone)
      const float *xb = in + ((size_t)b * T + min(t, T - 1)) * 64;
      float xr[64];
#pragma unroll
      for (int c4 = 0; c4 < 16; ++c4) {
        float4 v = __ldg(reinterpret_cast<const float4 *>(&xb[c4 * 4]));
        xr[c4 * 4] = v.x; xr[c4 * 4 + 1] = v.y; xr[c4 * 4 + 2] = v.z; xr[c4 * 4 + 3] = v.w;
      }
#pragma unroll
      for (int o = 0; o < WM_MAX_HEAD; ++o) {
        if (o < nout) {
          float a = bs[o];
#pragma unroll
          for (int c = 0; c < 64; ++c) a = fmaf(xr[c], ws[o * 64 + c], a);
          lg[o] = a;
        }
      }
    } else {
      if (t < T) {
        const float *lp = in + ((size_t)b * T + t) * nout;
#pragma unroll
        for (int o = 0; o < WM_MAX_HEAD; ++o)
          if (o < nout) lg[o] = lp[o];
      }
    }
    if (t < T) {
      float p = sigmoid_acc(lg[0]);
      if (probs) probs[(size_t)b * T + t] = p;
      if (t < valid) {
        sp += p;
#pragma unroll
        for (int o = 0; o < WM_MAX_HEAD - 1; ++o) {
          if (o < nbits) { sl[o] += lg[o + 1]; sv[o] += (lg[o + 1] > 0.0f) ? 1.0f : 0.0f; }
        }
      }
    }
  }
  const float inv = valid > 0 ? 1.0f / (float)valid : 0.0f;
  {
    float v = warp_sum(sp);
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (tid == 0 && clip_prob) {
      float s = 0.f;
      for (int i = 0; i < 8; ++i) s += red[i];
      clip_prob[b] = s * inv;
    }
  }
#pragma unroll
  for (int o = 0; o < WM_MAX_HEAD - 1; ++o) {
    if (o < nbits) {
      float v = warp_sum(sl[o]);
      float u = warp_sum(sv[o]);
      __syncthreads();
      if (lane == 0) red[warp] = v;
      __syncthreads();
      if (tid == 0 && msg_logits) {
        float s = 0.f;
        for (int i = 0; i < 8; ++i) s += red[i];
        msg_logits[(size_t)b * nbits + o] = s * inv;
      }
      __syncthreads();
      if (lane == 0) red[warp] = u;
      __syncthreads();
      if (tid == 0 && vote_frac) {
        float s = 0.f;
        for (int i = 0; i < 8; ++i) s += red[i];
        vote_frac[(size_t)b * nbits + o] = s * inv;
      }
    }
  }
}

int launch_detect_heads(const float *logits, const int *valid_len, float *probs, float *clip_prob,
                        float *msg_logits, float *vote_frac, int B, int T, int nout,
                        cudaStream_t st) {
  if (B == 0 || T == 0) return 0;
  if (nout < 1 || nout > WM_MAX_HEAD) { set_error("detect_heads: nout must be in [1,%d]", WM_MAX_HEAD); return -1; }
  detect_heads_kernel<false><<<B, 256, 0, st>>>(logits, nullptr, nullptr, valid_len, probs, clip_prob,
                                                msg_logits, vote_frac, T, nout);
  WM_CHECK_LAUNCH("detect_heads");
  return 0;
}

int launch_head_detect(const float *x, const float *w, const float *b, const int *valid_len,
                       float *probs, float *clip_prob, float *msg_logits, float *vote_frac, int B,
                       int T, int nout, cudaStream_t st) {
  if (B == 0 || T == 0) return 0;
  if (nout < 1 || nout > WM_MAX_HEAD) { set_error("head_detect: nout must be in [1,%d]", WM_MAX_HEAD); return -1; }
  detect_heads_kernel<true><<<B, 256, 0, st>>>(x, w, b, valid_len, probs, clip_prob, msg_logits,
                                               vote_frac, T, nout);
  WM_CHECK_LAUNCH("head_detect");
  return 0;
}

}  // namespace wm
