// Shared helpers for the wmb200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/wmb200.h"

namespace wm {

void set_error(const char *fmt, ...);
void count_launch(int n = 1);

#define WM_CHECK_ARG(cond, ...)          \
  do {                                   \
    if (!(cond)) {                       \
      wm::set_error(__VA_ARGS__);        \
      return -1;                         \
    }                                    \
  } while (0)

#define WM_CHECK_LAUNCH(name)                                                        \
  do {                                                                               \
    cudaError_t e_ = cudaGetLastError();                                             \
    if (e_ != cudaSuccess) {                                                         \
      wm::set_error("%s launch failed: %s", name, cudaGetErrorString(e_));           \
      return -2;                                                                     \
    }                                                                                \
    wm::count_launch();                                                              \
  } while (0)

#define WM_CHECK_CUDA(expr)                                                          \
  do {                                                                               \
    cudaError_t e_ = (expr);                                                         \
    if (e_ != cudaSuccess) {                                                         \
      wm::set_error("%s failed: %s", #expr, cudaGetErrorString(e_));                 \
      return -3;                                                                     \
    }                                                                                \
  } while (0)

#define WM_TRY(expr)        \
  do {                      \
    int rc_ = (expr);       \
    if (rc_ != 0) return rc_; \
  } while (0)

inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }
int sm_count();
int require_device();   // 0 when the current device is sm_10x, else sets the error and returns <0

__device__ __forceinline__ float sigmoid_acc(float x) {
  return __fdividef(1.0f, 1.0f + __expf(-x));
}
__device__ __forceinline__ float tanh_acc(float x) {
  // 2*sigmoid(2x) - 1 ; saturates cleanly for |x| large
  return __fdividef(2.0f, 1.0f + __expf(-2.0f * x)) - 1.0f;
}

// internal launchers shared between translation units (all async on `st`)
int launch_conv_in_k7(const float *s, const float *w, const float *b, float *y, int B, int T,
                      cudaStream_t st);
int launch_conv64_fp32(const float *x, const float *w, const float *bias, const float *residual,
                       const float *chan_add, float *y, int B, int T, int taps, int relu,
                       cudaStream_t st);
int launch_lstm_fp32(const float *x, const float *w_ih, const float *w_hh, const float *bias,
                     float *h, int B, int T, cudaStream_t st);
int launch_head(const float *x, const float *w, const float *b, float *y, int B, int T, int nout,
                cudaStream_t st);
int launch_gather_rows(const float *table, int64_t rows, const int64_t *idx, float *out, int B,
                       cudaStream_t st);
int launch_postprocess(const float *delta_raw, const float *s, const float *fir, float *delta,
                       float *s_w, float *rms_out, int B, int T, int mode, float peak,
                       float max_rms, float eps, cudaStream_t st);
int launch_detect_heads(const float *logits, const int *valid_len, float *probs, float *clip_prob,
                        float *msg_logits, float *vote_frac, int B, int T, int nout,
                        cudaStream_t st);
// head (64 -> nout) + sigmoid + per-clip reductions straight from the last activation
int launch_head_detect(const float *x, const float *w, const float *b, const int *valid_len,
                       float *probs, float *clip_prob, float *msg_logits, float *vote_frac, int B,
                       int T, int nout, cudaStream_t st);

// tcgen05 path (wm_conv_tc.cu); x / residual / y are planar activations, w_img from launch_pack_conv64_tc
int launch_conv64_tc(const void *x, const void *w_img, const float *bias, const void *residual, void *y, float *y32,
                     int B, int T, int taps, int relu, cudaStream_t st, const float *res32 = nullptr);
int launch_pack_conv64_tc(const float *w, void *img, int taps, cudaStream_t st);
int launch_to_planar(const float *x, const float *chan_add, void *y, int B, int T, cudaStream_t st);
int launch_from_planar(const void *x, float *y, int B, int T, cudaStream_t st);
int launch_conv_in_k7_planar(const float *s, const float *w, const float *b, void *y, int B, int T, cudaStream_t st);
int launch_lstm_tc(const void *x, const void *wpk, const float *bias_p, const float *chan_add, void *y, int B, int T,
                   cudaStream_t st);
// host_b12: HOST copy of b1[64] then b2[64] (biases become constant-bank operands), or null (read from b1 / b2)
int launch_resblock_tc(const void *x, const void *w_img, const float *b1, const float *b2, void *y, float *y32, int B,
                       int T, cudaStream_t st, const float *host_b12 = nullptr);
int resblock_tiles_per_clip(int T);
// first ResBlock fused with the input convolution (wm_resblock_in_tc.cu); w9b: device w9[9][64], b9[64];
// winb: device w_in[7][64], b_in[64]; fin: device WM_FIN_* block; s[B][T] -> y planar
int launch_resblock_in_tc(const float *s, const float *w9b, const float *winb, const float *fin, const void *w_img2,
                          const float *b2, void *y, int B, int T, cudaStream_t st);
// host_head: HOST copy of output 0 of the 1x1 head (w[64] then b); it is passed to the kernel by value
int launch_resblock_head1_tc(const void *x, const void *w_img, const float *b1, const float *b2,
                             const float *host_head, float *delta_raw, int B, int T, cudaStream_t st,
                             const float *host_b12 = nullptr);
#define WM_DET_PART 84   // floats per (tile, warp) partial of the detector epilogue: 64 activation sums, prob sum, pad, 16 vote counts
#define WM_DET_VOTE0 68  // first vote count inside a partial
#define WM_FUSED_VOTE_MAX 17   // the fused majority-vote epilogue handles heads of up to 1 + 16 outputs (py/main16.py:34)
int launch_resblock_detect_tc(const void *x, const void *w_img, const float *b1, const float *b2,
                              const float *host_head, const int *valid_len, float *probs, float *partials, int B, int T,
                              cudaStream_t st, const float *host_b12 = nullptr, const float *host_head_all = nullptr,
                              int nout = 0);
int launch_detect_finalize(const float *partials, const int *valid_len, const float *head_w, const float *head_b,
                           float *clip_prob, float *msg_logits, float *vote_frac, int B, int T, int nout, cudaStream_t st);
// training-loss forward kernels (wm_loss.cu); `partials` is caller workspace (wm_loss_workspace_bytes)
int launch_stft_mag(const float *x, float *mag, int B, int T, int n_fft, int hop, cudaStream_t st);
int launch_hf_penalty(const float *delta, float *out, float *partials, int B, int T, int n_fft, int first_bin,
                      cudaStream_t st);
int launch_loudness(const float *clean, const float *wmk, float *out, float *partials, int B, int T, int n_fft, int hop,
                    float thresh, cudaStream_t st);
int launch_mel_log_l1(const float *clean, const float *wmk, const float *fb, const int *band, int n_mels, float *out,
                      float *partials, int B, int T, int n_fft, int hop, cudaStream_t st);
int launch_bce_heads(const float *logits, const int64_t *message, float *loc_out, float *bce_out, float *partials,
                     int B_wm, int B2, int T, int nout, cudaStream_t st);
int launch_abs_mean(const float *x, long long n, float *out, float *partials, cudaStream_t st);
size_t stft_bwd_scratch_floats(int B, int T, int n_fft, int hop);
int launch_hf_penalty_bwd(const float *delta, float *d_delta, float *gframes, int B, int T, int n_fft, int first_bin,
                          float weight, int accumulate, cudaStream_t st);
int launch_loudness_bwd(const float *clean, const float *wmk, float *d_wmk, float *gframes, int B, int T, int n_fft,
                        int hop, float thresh, float weight, int accumulate, cudaStream_t st);
int launch_mel_log_l1_bwd(const float *clean, const float *wmk, const float *fb, const int *band, int n_mels,
                          float *d_wmk, float *gframes, int B, int T, int n_fft, int hop, float weight, int accumulate,
                          cudaStream_t st);
int launch_abs_mean_bwd(const float *x, float *dx, long long n, float weight, int accumulate, cudaStream_t st);
int launch_postprocess_bwd(const float *g, const float *d1, const float *fir, float *d_delta_raw, float *scratch, int B,
                           int T, int mode, float peak, float max_rms, float eps, cudaStream_t st);
// training building blocks (wm_train.cu)
size_t train_scratch_doubles(long long N);
size_t conv_wgrad_scratch_floats(int B, int T, int K);
size_t detector_train_workspace_bytes(int B2, int T, int nout);
int launch_bn_train_fwd(const float *z, const float *gamma, const float *beta, const float *residual, float *out,
                        float *mean, float *rstd, float *run_mean, float *run_var, long long N, int relu, double *scratch,
                        cudaStream_t st);
int launch_bn_train_bwd(const float *dout, const float *act, const float *z, const float *mean, const float *rstd,
                        const float *gamma, float *dz, float *dres, float *dgamma, float *dbeta, long long N,
                        double *scratch, cudaStream_t st);
int launch_transpose_flip(const float *w, float *wt, int K, cudaStream_t st);
int launch_bce_heads_bwd(const float *logits, const int64_t *message, int B_wm, int B2, int T, int nout, float lam_loc,
                         float lam_dec, float *dlog, cudaStream_t st);
size_t head_bwd_scratch_floats(long long N, int nout);
int launch_head_bwd(const float *dlog, const float *y, const float *w, float *dy, float *dw, float *db, long long N,
                    int nout, float *scratch, cudaStream_t st);
size_t conv_in_grads_scratch_floats(int B, int T);
int launch_conv_in_grads(const float *s, const float *dx, const float *w, float *dw, float *db, float *ds, int B, int T,
                         float *scratch, cudaStream_t st);
int launch_conv_wgrad(const float *x, const float *dz, float *dw, float *db, int B, int T, int K, float *scratch,
                      cudaStream_t st);
int launch_conv_wgrad_ex(const float *x, const float *dz, float *dw, float *db, int B, int T, int K, int P, int bias_tap,
                         float *scratch, cudaStream_t st);
// training LSTM (wm_train_lstm.cu); weights per-gate transposed wT[q][k][r] = W[q*64 + r][k]
int launch_lstm_train_fwd(const float *x, const float *wT_ih, const float *wT_hh, const float *b_ih, const float *b_hh,
                          float *h, float *gates, float *cell, int B, int T, cudaStream_t st);
size_t lstm_train_bwd_scratch_floats(int B, int T);
int launch_lstm_train_bwd(const float *dy, const float *x, const float *h, const float *wT_ih, const float *wT_hh,
                          const float *gates, const float *cell, float *dx, float *dwT_ih, float *dwT_hh, float *db,
                          int B, int T, float *scratch, cudaStream_t st);
size_t train_step_workspace_bytes(int B, int T, int nout);
int train_forward_backward(const float *g_params, float *g_grads, float *g_stats, const float *d_params, float *d_grads,
                           float *d_stats, const float *s, const int64_t *message, const float *fir, const float *mel_fb,
                           const int *mel_band, int n_mels, const float *lam, int B, int T, int nout, float *losses_out,
                           float *s_w_out, void *workspace, cudaStream_t st);
int launch_adam(float *p, const float *g, float *m, float *v, long long n, float lr, float b1, float b2, float eps,
                int step, cudaStream_t st);
int detector_train_step(float *params, float *grads, float *adam_m, float *adam_v, float *run_stats, const float *x,
                        const int64_t *message, int B_wm, int B2, int T, int nout, float lam_loc, float lam_dec, float lr,
                        float beta1, float beta2, float eps, int adam_step, float *losses_out, float *d_input,
                        void *workspace, cudaStream_t st);
// generic fp32 operators of the main14b_2 stack (wm_generic.cu); channels-first x[b][c][t]
int launch_conv1d_generic(const float *x, const float *w, const float *bias, const float *chan_add, const float *res,
                          float *y, int B, int Cin, int Tin, int Cout, int K, int stride, int pad, int act,
                          cudaStream_t st, int shuffle = 1, int Tstore = 0, int extra_out = 0);
int launch_convt_phase_weights(const float *w, const float *bias, float *w3, float *b3, int Cin, int Cout, int s, int p,
                               cudaStream_t st);
int launch_convtranspose1d_generic(const float *x, const float *w, const float *bias, float *y, int B, int Cin, int Tin,
                                   int Cout, int K, int stride, int pad, cudaStream_t st);
int launch_lstm_small(const float *x, const float *w_ih, const float *w_hh, const float *bias, float *y, int B, int H,
                      int T, int L, cudaStream_t st);
// wm_m14_small.cu: 0 = done, 1 = shape not handled there (caller uses the generic kernel), < 0 error
int launch_lstm_small_reg(const float *x, const float *w_ih, const float *w_hh, const float *bias, float *y, int B, int H,
                          int T, int L, cudaStream_t st);
// audio formats either side of the path (wm_audio.cu)
int launch_resample(const float *x, const float *kern, float *y, int B, int Tin, int Tout, int down, int up, int K,
                    int width, cudaStream_t st);
int launch_pcm16(const float *x, short *q, float *xo, long long n, int quantize, float scale, cudaStream_t st);
int launch_file_metrics(const float *s, const float *sw, const int *valid_len, float *out, int B, int T, cudaStream_t st);
// wm_wgrad_tc.cu: tcgen05 weight gradient of the 64 -> 64 convolutions (planar operands), bias gradient
size_t wgrad_tc_scratch_floats(int K);
int launch_wgrad_tc(const void *xp, const void *dyp, float *dw, int B, int T, int K, float *scratch, cudaStream_t st);
size_t colsum64_scratch_floats();
int launch_colsum64(const float *dy, float *db, long long rows, float *scratch, cudaStream_t st);
int math_mode();   // WM_MATH_FP32 / WM_MATH_BF16X2 (wm_set_math_mode)
// wm_train.cu: the training step's 64 -> 64 convolution as single operators (tensor cores in the default math mode)
size_t train_conv64_scratch_floats(int B, int T, int K);
int train_conv64_fwd(const float *x, const float *w, const float *bias, const float *residual, float *y, int B, int T,
                     int K, float *scratch, cudaStream_t st);
int train_conv64_bwd(const float *x, const float *dy, const float *w, float *dw, float *db, float *dx, int B, int T, int K,
                     float *scratch, cudaStream_t st);
// wm_eval.cu: second-order IIR (lfilter semantics) + PCM16, detection statistics
size_t biquad_scratch_bytes(int rows, long long N);
int launch_biquad(const float *x, float *y, short *q, int rows, long long N, const double *b, const double *a, int clamp,
                  void *scratch, cudaStream_t st);
int launch_confusion(const float *clean, long long n0, const float *wm, long long n1, float thresh, unsigned long long *out4,
                     cudaStream_t st);
int launch_roc_points(const float *clean, long long n0, const float *wm, long long n1, const float *thr, int nt, int *fp,
                      int *tp, cudaStream_t st);
int launch_auc_pairs(const float *clean, long long n0, const float *wm, long long n1, unsigned long long *out, cudaStream_t st);
void set_lstm_profile_buffer(long long *p);
void set_lstm_opts(int opts);
int get_debug_opts();   // developer A/B switches (wm_debug_lstm_opts): bit 8 = pconv without stores, bit 9 = pconv without MMAs
long long *get_profile_buffer();
int launch_pack_lstm_tc(const float *w_ih, const float *w_hh, const float *bias, void *wpk, float *bias_p,
                        cudaStream_t st);

}  // namespace wm
