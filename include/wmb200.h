/*
 * wmb200 — C ABI of the B200-native main16 embed+detect path.
 *
 * The reference (Spandan7724/Audio-Watermarking-...) is pure Python and has no
 * FFI; every entry point below replaces the arithmetic of one Python call site
 * of the reference (file:line given per function, relative to the reference
 * root).  A maintainer of the reference binds this library with ctypes exactly
 * as `audio-watermarking-..._b200/_lib.py` does.
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless the name says `host`;
 *   - waveforms are fp32 [b][t]; hidden activations of the single-operator
 *     entry points are fp32 channels-last x[b][t][c], c = 0..63 contiguous;
 *   - every function is asynchronous on `stream` (a cudaStream_t passed as
 *     void*), never allocates device memory and never synchronises, with two
 *     documented exceptions: wm_finalize_generator_blob / wm_finalize_detector_blob
 *     (one-time set-up per weight blob) read the blob's fp32 block back to the host
 *     and wait for that copy -- the fused kernels take biases and 1x1-head weights by
 *     value (constant bank), which needs them on the host; and wm_embed_detect_host
 *     keeps two copy streams and a handful of events per calling host thread
 *     (created on first use, never destroyed) to overlap H2D / D2H with the kernels:
 *     `stream` is made to wait for every one of them before the call returns, so
 *     stream-order semantics are those of a plain sequence on `stream`;
 *   - return value 0 = success, negative = error; wm_last_error() returns the
 *     message of the last failure on the calling thread;
 *   - weights arrive as one packed fp32 blob per module (layout below), built
 *     on the host side from the reference state dict (eval BatchNorm folded).
 *   - the library only runs on compute capability 10.x; there is no fallback.
 */
#ifndef WMB200_H
#define WMB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WM_C 64            /* channels of every hidden activation (py/main16.py:134) */
#define WM_FIR_TAPS 101    /* py/main16.py:53 */
#define WM_MAX_HEAD 32     /* max outputs of the 1x1 head (1 + message_bits)        */
#define WM_ABI_VERSION 21
#define WM_PLANAR_PAD 4      /* zero rows before / after every plane of the planar layout */
#define WM_POST_FIR 1
#define WM_POST_CLAMP 2
#define WM_POST_RMS 4
#define WM_POST_ALL 7

/* ---- packed fp32 weight blobs (offsets in floats) ------------------------
 * conv weights are stored tap-major:  w[j][ci][co]  (co contiguous), i.e. the
 * reference's Conv1d weight (co,ci,j) permuted to (j,ci,co) with the eval-mode
 * BatchNorm scale folded in; ConvTranspose1d (ci,co,j) is flipped along j and
 * stored in the same (j,ci,co) form.  Biases have BatchNorm folded in.       */
enum {
  WM_RB_W1 = 0,                          /* [3][64][64] */
  WM_RB_B1 = WM_RB_W1 + 3 * 64 * 64,     /* [64]        */
  WM_RB_W2 = WM_RB_B1 + 64,              /* [3][64][64] */
  WM_RB_B2 = WM_RB_W2 + 3 * 64 * 64,     /* [64]        */
  WM_RB_SIZE = WM_RB_B2 + 64
};
/* Input stage fused into the first ResBlock (wm_resblock_in_tc.cu): the input Conv1d(1,64,7) composed with
 * the ResBlock's first convolution on the host, in float64:  conv1(conv_in(s))[t] = B9 + sum_m W9[m] s[t+m-4]. */
enum {
  WM_FIN_W9 = 0,                         /* [9][64]    W9[k+j][co] += sum_ci W1[k][ci][co] w_in[j][ci]            */
  WM_FIN_B9 = WM_FIN_W9 + 9 * 64,        /* [64]       b1 + sum_k BK[k]                                            */
  WM_FIN_WK = WM_FIN_B9 + 64,            /* [3][7][64] per conv1 tap k: sum_ci W1[k][ci][co] w_in[j][ci]           */
  WM_FIN_BK = WM_FIN_WK + 3 * 7 * 64,    /* [3][64]    per conv1 tap k: sum_ci W1[k][ci][co] b_in[ci]              */
  WM_FIN_SIZE = WM_FIN_BK + 3 * 64
};
enum {                                   /* Generator, py/main16.py:128-162 */
  WM_G_IN_W = 0,                         /* encoder.0  [7][64]       */
  WM_G_IN_B = WM_G_IN_W + 7 * 64,        /* [64]                     */
  WM_G_RB0 = WM_G_IN_B + 64,             /* encoder.1                */
  WM_G_RB1 = WM_G_RB0 + WM_RB_SIZE,      /* encoder.2                */
  WM_G_LSTM_WIH = WM_G_RB1 + WM_RB_SIZE, /* [256][64] rows i,f,g,o   */
  WM_G_LSTM_WHH = WM_G_LSTM_WIH + 256 * 64,
  WM_G_LSTM_B = WM_G_LSTM_WHH + 256 * 64,/* [256] = b_ih + b_hh      */
  WM_G_CT_W = WM_G_LSTM_B + 256,         /* decoder.0 as conv [7][64][64] */
  WM_G_CT_B = WM_G_CT_W + 7 * 64 * 64,   /* [64]                     */
  WM_G_RB2 = WM_G_CT_B + 64,             /* decoder.1                */
  WM_G_HEAD_W = WM_G_RB2 + WM_RB_SIZE,   /* decoder.2 [64]           */
  WM_G_HEAD_B = WM_G_HEAD_W + 64,        /* [1] (+3 pad)             */
  WM_G_FIN = WM_G_HEAD_B + 4,            /* encoder.0 composed with encoder.1's first convolution */
  WM_G_SIZE = WM_G_FIN + WM_FIN_SIZE,
  /* tcgen05 weight images (bf16 [tap][ci/8][128][8], see wm_pack_conv64_tc), filled on the
   * device by wm_finalize_generator_blob: encoder.1 conv1, conv2, encoder.2 conv1, conv2,
   * decoder.0 (7 taps), decoder.1 conv1, conv2 */
  WM_G_TC = (WM_G_SIZE + 63) / 64 * 64,
  WM_TC_IMG3 = 3 * 8 * 128 * 8 / 2,      /* floats occupied by a 3-tap image */
  WM_TC_IMG7 = 7 * 8 * 128 * 8 / 2,
  WM_G_TC_CT = WM_G_TC + 4 * WM_TC_IMG3,
  WM_G_TC_RB2 = WM_G_TC_CT + WM_TC_IMG7,
  /* tensor-core LSTM operands (wm_pack_lstm_tc): bf16 [4][2][128][64] then fp32 bias [2][128] */
  WM_G_TC_LSTM_W = WM_G_TC_RB2 + 2 * WM_TC_IMG3,
  WM_G_TC_LSTM_B = WM_G_TC_LSTM_W + 4 * 256 * 64 / 2,
  WM_G_BLOB = WM_G_TC_LSTM_B + 256
};
enum {                                   /* Detector, py/main16.py:170-186 */
  WM_D_IN_W = 0,                         /* model.0 [7][64]          */
  WM_D_IN_B = WM_D_IN_W + 7 * 64,
  WM_D_RB0 = WM_D_IN_B + 64,             /* model.1                  */
  WM_D_RB1 = WM_D_RB0 + WM_RB_SIZE,      /* model.2                  */
  WM_D_HEAD_W = WM_D_RB1 + WM_RB_SIZE,   /* model.3 [nout<=32][64]   */
  WM_D_HEAD_B = WM_D_HEAD_W + 32 * 64,   /* [32]                     */
  WM_D_FIN = WM_D_HEAD_B + 32,           /* model.0 composed with model.1's first convolution */
  WM_D_SIZE = WM_D_FIN + WM_FIN_SIZE,
  WM_D_TC = (WM_D_SIZE + 63) / 64 * 64,  /* model.1 conv1, conv2, model.2 conv1, conv2 */
  WM_D_BLOB = WM_D_TC + 4 * WM_TC_IMG3
};

/* Training-mode Detector parameters (BatchNorm NOT folded), one flat fp32 buffer that Adam updates in place;
 * gradients and both Adam moments use the same layout.  Convolution weights are tap-major [k][ci][co]. */
enum {
  WM_DT_RB_W1 = 0,                         /* conv1.weight [3][64][64]                  */
  WM_DT_RB_B1 = WM_DT_RB_W1 + 3 * 64 * 64, /* conv1.bias                                 */
  WM_DT_RB_G1 = WM_DT_RB_B1 + 64,          /* bn1.weight                                 */
  WM_DT_RB_BE1 = WM_DT_RB_G1 + 64,         /* bn1.bias                                   */
  WM_DT_RB_W2 = WM_DT_RB_BE1 + 64,
  WM_DT_RB_B2 = WM_DT_RB_W2 + 3 * 64 * 64,
  WM_DT_RB_G2 = WM_DT_RB_B2 + 64,
  WM_DT_RB_BE2 = WM_DT_RB_G2 + 64,
  WM_DT_RB_SIZE = WM_DT_RB_BE2 + 64,
  WM_DT_IN_W = 0,                          /* model.0 [7][64]                            */
  WM_DT_IN_B = WM_DT_IN_W + 7 * 64,
  WM_DT_RB0 = WM_DT_IN_B + 64,             /* model.1, model.2                           */
  WM_DT_HEAD_W = WM_DT_RB0 + 2 * WM_DT_RB_SIZE, /* model.3 [32][64] (rows >= nout zero)  */
  WM_DT_HEAD_B = WM_DT_HEAD_W + 32 * 64,
  WM_DT_SIZE = WM_DT_HEAD_B + 32,
  WM_DT_STATS = 2 * 4 * 64                 /* per ResBlock: bn1 running_mean, running_var, bn2 ... */
};

/* Training-mode Generator parameters, same conventions (ResBlocks use the WM_DT_RB_* block; LSTM weights per-gate
 * transposed wT[q][k][r] = W[q*64 + r][k]; decoder.0 as a convolution [7][ci][co] = weight[ci][co][6 - j]). */
enum {
  WM_GT_IN_W = 0,                           /* encoder.0 [7][64]                          */
  WM_GT_IN_B = WM_GT_IN_W + 7 * 64,
  WM_GT_RB0 = WM_GT_IN_B + 64,              /* encoder.1                                  */
  WM_GT_RB1 = WM_GT_RB0 + WM_DT_RB_SIZE,    /* encoder.2                                  */
  WM_GT_LSTM_WIH = WM_GT_RB1 + WM_DT_RB_SIZE,
  WM_GT_LSTM_WHH = WM_GT_LSTM_WIH + 256 * 64,
  WM_GT_LSTM_BIH = WM_GT_LSTM_WHH + 256 * 64,
  WM_GT_LSTM_BHH = WM_GT_LSTM_BIH + 256,
  WM_GT_CT_W = WM_GT_LSTM_BHH + 256,        /* decoder.0                                  */
  WM_GT_CT_B = WM_GT_CT_W + 7 * 64 * 64,
  WM_GT_RB2 = WM_GT_CT_B + 64,              /* decoder.1                                  */
  WM_GT_HEAD_W = WM_GT_RB2 + WM_DT_RB_SIZE, /* decoder.2 [64]                             */
  WM_GT_HEAD_B = WM_GT_HEAD_W + 64,         /* [1] (+63 pad)                              */
  WM_GT_EMB = WM_GT_HEAD_B + 64,            /* embedding.weight [65536][64]               */
  WM_GT_SIZE = WM_GT_EMB + 65536 * 64,
  WM_GT_STATS = 3 * 4 * 64
};

/* Arithmetic of the 64->64 convolutions (the tensor-pipe part of the path).
 *   WM_MATH_FP32   CUDA-core fp32 FMA (bit-for-bit the reference's operation order
 *                  up to summation order; the on-GPU cross-check for the other mode)
 *   WM_MATH_BF16X2 tcgen05 kind::f16 with every operand split into bf16 hi+lo and
 *                  all four partial products accumulated in fp32 TMEM              */
enum { WM_MATH_FP32 = 0, WM_MATH_BF16X2 = 1 };

int wm_abi_version(void);
const char *wm_last_error(void);
/* 1 when the current device is compute capability 10.x; the library refuses
 * to launch anywhere else (there is no fallback path). */
int wm_device_ok(void);
/* Select the conv arithmetic for subsequent calls of this process (default
 * WM_MATH_BF16X2 once available).  Returns the previous mode. */
int wm_set_math_mode(int mode);
int wm_get_math_mode(void);
/* Number of kernels this library has launched since load (bench.py gpu_launches). */
unsigned long long wm_launch_count(void);

/* The host packs the first WM_G_SIZE / WM_D_SIZE floats of a WM_G_BLOB / WM_D_BLOB float
 * device buffer; these fill the tcgen05 weight images behind them (one small kernel per conv). */
int wm_finalize_generator_blob(float *blob, void *stream);
int wm_finalize_detector_blob(float *blob, void *stream);

/* ---- single operators ---------------------------------------------------*/

/* nn.Conv1d(1,64,7,padding=3)  — py/main16.py:134 (Generator) and :177 (Detector).
 * s[B][T] -> y[B][T][64];  w[7][64], b[64]. */
int wm_conv_in_k7_fwd(const float *s, const float *w, const float *b, float *y,
                      int B, int T, void *stream);

/* One 64->64 convolution with fused epilogue — the Conv1d(64,64,3,p=1)+BatchNorm
 * (+ReLU) halves of ResBlock, py/main16.py:116-121, and ConvTranspose1d(64,64,7,
 * p=3), py/main16.py:144 (taps = 7, weights pre-flipped).
 *   y = act( conv(x + chan_add) + bias + residual )
 * chan_add (nullable) is a per-clip [B][64] vector added to every in-range time
 * step of x before the convolution (the message embedding, py/main16.py:156-159);
 * residual (nullable) is [B][T][64]; relu != 0 applies max(.,0). */
int wm_conv64_fwd(const float *x, const float *w, const float *bias, const float *residual,
                  const float *chan_add, float *y, int B, int T, int taps, int relu,
                  void *stream);

/* ---- the tensor-core form of the same convolution -------------------------------------
 * "planar" activations: per clip 16 planes of (T + 2*WM_PLANAR_PAD) rows x 16 bytes; plane c
 * (0..7) holds the bf16 high parts of channels 8c..8c+7 of every time step, plane 8+c the bf16
 * low parts (v = hi + lo), with WM_PLANAR_PAD zero rows before t = 0 and after t = T-1.       */
size_t wm_planar_bytes(int B, int T);
/* fp32 channels-last x[B][T][64] (+ chan_add[B][64], nullable) -> planar */
int wm_to_planar(const float *x, const float *chan_add, void *y, int B, int T, void *stream);
int wm_from_planar(const void *x, float *y, int B, int T, void *stream);
/* fp32 w[taps][64][64] (tap-major, as in the blobs) -> bf16 image of wm_conv64_tc_weight_bytes(taps) */
size_t wm_conv64_tc_weight_bytes(int taps);
int wm_pack_conv64_tc(const float *w, void *img, int taps, void *stream);
/* y = act(conv(x) + bias + residual) on planar tensors; taps 3 or 7; y (planar) and y32
 * (fp32 channels-last [B][T][64]) are both optional outputs.  tcgen05.mma kind::f16, every
 * operand a bf16 hi+lo pair, all four partial products accumulated in fp32 (TMEM). */
int wm_conv64_tc_fwd(const void *x, const void *w_img, const float *bias, const void *residual,
                     void *y, float *y32, int B, int T, int taps, int relu, void *stream);

/* A whole ResBlock (py/main16.py:112-125) in one kernel: y = relu(x + conv2(relu(conv1(x)+b1)) + b2).
 * w_img: the two 3-tap images (conv1 then conv2) back to back; the intermediate activation stays
 * in shared memory, the residual is taken from the x tile already on chip. */
int wm_resblock_tc_fwd(const void *x, const void *w_img, const float *b1, const float *b2, void *y,
                       float *y32, int B, int T, void *stream);
/* The same with the two bias vectors read from HOST memory at call time and passed to the kernel by value
 * (constant-bank operands: the variant the module-level drivers use once a blob has been finalized). */
int wm_resblock_tc_hostbias_fwd(const void *x, const void *w_img, const float *host_b1, const float *host_b2,
                                void *y, float *y32, int B, int T, void *stream);

/* The LSTM on tensor cores: planar x -> planar h (+ chan_add[B][64] added to the OUTPUT only, i.e.
 * the message embedding of py/main16.py:156-159 fused into the store).  Weights resident in TMEM,
 * bf16 hi+lo operand pairs, fp32 accumulation and fp32 cell state.  wpk / bias_p from wm_pack_lstm_tc
 * (wpk: 4*256*64 bf16, bias_p: 256 floats). */
int wm_pack_lstm_tc(const float *w_ih, const float *w_hh, const float *bias, void *wpk, float *bias_p,
                    void *stream);
/* developer hook: when buf16 (device, 16 x int64) is non-null the LSTM kernel's block 0 stores
 * per-phase cycle sums there (tools/lstm_profile.py); pass NULL to switch it off. */
int wm_debug_lstm_profile(long long *buf16);
/* developer hook: scheduling variants of the LSTM kernel for A/B measurements (tools/lstm_profile.py);
 * bit 0: the two clip groups of a CTA take turns on the MUFU-heavy part of a step, bit 1: generic->async proxy
 * fence issued by the MMA thread instead of by every epilogue thread, bit 2: accumulator released after the
 * exponentials.  -1 restores the library default.  Results are identical in every mode. */
int wm_debug_lstm_opts(int opts);
int wm_lstm_tc_fwd(const void *x, const void *wpk, const float *bias_p, const float *chan_add, void *y,
                   int B, int T, void *stream);

/* nn.LSTM(64,64,batch_first=True) with zero initial state, all hidden states
 * returned — py/main16.py:138,153.  x[B][T][64] -> h[B][T][64].
 * w_ih, w_hh [256][64] (rows i,f,g,o), bias[256] = b_ih + b_hh. */
int wm_lstm_fwd(const float *x, const float *w_ih, const float *w_hh, const float *bias,
                float *h, int B, int T, void *stream);

/* Conv1d(64,nout,1): py/main16.py:146 (nout = 1) and :180 (nout = 1+bits <= 32).
 * x[B][T][64] -> y[B][T][nout] (this IS the permute(0,2,1) view of :186). */
int wm_head_fwd(const float *x, const float *w, const float *b, float *y,
                int B, int T, int nout, void *stream);

/* Delta post-processing + mix, py/main16.py:245-248 (fir_lowpass :53-64,
 * clamp_peak :66-67, limit_rms :69-72, s + delta :248).
 * mode is a bit mask: WM_POST_FIR | WM_POST_CLAMP | WM_POST_RMS.
 *   mode 7: delta = limit_rms(clamp_peak(fir(delta_raw)));  s_w = s + delta   (:245-248)
 *   mode 0: delta = delta_raw;  s_w = s + delta   (generate_watermarked_audio, :1006)
 * fir[101] are the taps (device, nullable unless WM_POST_FIR); delta / s_w nullable;
 * rms_out (nullable) [B] receives sqrt(mean(delta^2)) of the final delta
 * (evaluate_model, py/main16.py:403). */
int wm_postprocess_fwd(const float *delta_raw, const float *s, const float *fir,
                       float *delta, float *s_w, float *rms_out, int B, int T, int mode,
                       float peak, float max_rms, float eps, void *stream);

/* Detection heads on logits[B][T][nout] — py/main16.py:1142-1146, :393-398.
 *   probs[B][T]          sigmoid(logits[:,:,0])                    (nullable)
 *   clip_prob[B]         mean over the first valid_len[b] samples
 *   msg_logits[B][nout-1] mean over valid samples of logits[:,:,1:]
 *   vote_frac[B][nout-1]  fraction of valid samples with logit > 0 (nullable)
 * valid_len (nullable, int32 [B]) defaults to T (tail segments, :1159-1164). */
int wm_detect_heads_fwd(const float *logits, const int *valid_len, float *probs,
                        float *clip_prob, float *msg_logits, float *vote_frac,
                        int B, int T, int nout, void *stream);

/* ---- module-level drivers ----------------------------------------------*/

size_t wm_generator_workspace_bytes(int B, int T);
/* Generator.forward, py/main16.py:149-162.  s[B][T], message (nullable, int64 [B]),
 * embedding (nullable) [emb_rows][64] -> delta_raw[B][T]. */
int wm_generator_fwd(const float *blob, const float *embedding, int64_t emb_rows,
                     const int64_t *message, const float *s, float *delta_raw,
                     void *workspace, size_t workspace_bytes, int B, int T, void *stream);

size_t wm_detector_workspace_bytes(int B, int T);
/* Detector.forward, py/main16.py:183-186.  x[B][T] -> logits[B][T][nout]. */
int wm_detector_fwd(const float *blob, const float *x, float *logits, void *workspace,
                    size_t workspace_bytes, int B, int T, int nout, void *stream);

/* Detector + heads without materialising the logits: the per-segment body of
 * detect_watermark, py/main16.py:1140-1146 (and evaluate_model :392-398).
 * Outputs as wm_detect_heads_fwd. */
int wm_detect_fwd(const float *blob, const float *x, const int *valid_len, float *probs,
                  float *clip_prob, float *msg_logits, float *vote_frac, void *workspace,
                  size_t workspace_bytes, int B, int T, int nout, void *stream);

/* The benchmark unit: one batch of clips through
 *   G -> fir/clamp/rms (post_mode WM_POST_ALL) or raw (post_mode 0) -> s + delta -> D -> heads,
 * i.e. the forward of py/main16.py:378-398.  All outputs nullable except s_w. */
size_t wm_embed_detect_workspace_bytes(int B, int T);
int wm_embed_detect_fwd(const float *g_blob, const float *embedding, int64_t emb_rows,
                        const float *d_blob, const float *fir, const int64_t *message,
                        const float *s, float *delta, float *s_w, float *delta_rms,
                        float *probs, float *clip_prob, float *msg_logits, float *vote_frac,
                        void *workspace, size_t workspace_bytes, int B, int T, int nout,
                        int post_mode, void *stream);

/* ---- training-loss forward -------------------------------------------------
 * Scalars land in device memory (out[0]); `workspace` holds per-CTA partial sums that a
 * single-block kernel adds in a fixed order, so every loss is deterministic.            */
size_t wm_loss_workspace_bytes(int B, int T);
/* frames of torch.stft(centre=True): 1 + T / hop */
int wm_stft_frames(int T, int hop);
/* torch.stft(x, n_fft, hop, window=hann_window(n_fft), return_complex=True).abs()
 * (py/main16.py:77, 212-213): x[B][T] -> mag[B][n_fft/2+1][frames]; n_fft in {512,1024,2048},
 * reflect padding needs T > n_fft/2. */
int wm_stft_mag_fwd(const float *x, float *mag, int B, int T, int n_fft, int hop, void *stream);
/* high_freq_penalty — py/main16.py:74-81: mean over (B, n_fft/2+1, frames) of |STFT(delta)| on bins
 * >= first_bin (hop = n_fft/4; the reference's cutoff 3500 Hz at n_fft 512 is first_bin 113). */
int wm_hf_penalty_fwd(const float *delta, float *out, void *workspace, size_t workspace_bytes, int B, int T,
                      int n_fft, int first_bin, void *stream);
/* TFLoudnessLoss.forward — py/main16.py:204-217 (n_fft 2048, hop 512, thresh 0.01, strict >). */
int wm_loud_fwd(const float *clean, const float *watermarked, float *out, void *workspace, size_t workspace_bytes,
                int B, int T, int n_fft, int hop, float thresh, void *stream);
/* MultiScaleMelLoss.forward — py/main16.py:192-202: fb[n_fft/2+1][n_mels] is the torchaudio HTK
 * filterbank, band[2m], band[2m+1] the half-open bin range on which filter m is non-zero. */
int wm_mel_log_l1_fwd(const float *clean, const float *watermarked, const float *fb, const int *band, int n_mels,
                      float *out, void *workspace, size_t workspace_bytes, int B, int T, int n_fft, int hop,
                      void *stream);
/* F.binary_cross_entropy_with_logits of py/main16.py:255-264 on logits[B_total][T][nout]:
 * loc_out = mean over all rows of BCE(channel 0, [clip < B_wm]); bce_out (nullable) = mean over the
 * first B_wm clips of BCE(channel 1+j, bit j of message[b]). */
int wm_bce_heads_fwd(const float *logits, const int64_t *message, float *loc_out, float *bce_out, void *workspace,
                     size_t workspace_bytes, int B_wm, int B_total, int T, int nout, void *stream);
/* F.l1_loss(delta, 0) — py/main16.py:266. */
int wm_abs_mean_fwd(const float *x, float *out, void *workspace, size_t workspace_bytes, int B, int T, void *stream);

/* ---- backward of the losses and of the post-processing (py/main16.py:266-277 through autograd) ----
 * Each adds (accumulate != 0) or writes `weight` * dLoss/dSignal into a [B][T] gradient; `workspace` holds the
 * per-frame gradients (wm_stft_bwd_workspace_bytes) that a gather kernel overlap-adds, so results are deterministic. */
size_t wm_stft_bwd_workspace_bytes(int B, int T, int n_fft, int hop);
int wm_hf_penalty_bwd(const float *delta, float *d_delta, void *workspace, size_t workspace_bytes, int B, int T,
                      int n_fft, int first_bin, float weight, int accumulate, void *stream);
int wm_loud_bwd(const float *clean, const float *watermarked, float *d_watermarked, void *workspace,
                size_t workspace_bytes, int B, int T, int n_fft, int hop, float thresh, float weight, int accumulate,
                void *stream);
int wm_mel_log_l1_bwd(const float *clean, const float *watermarked, const float *fb, const int *band, int n_mels,
                      float *d_watermarked, void *workspace, size_t workspace_bytes, int B, int T, int n_fft, int hop,
                      float weight, int accumulate, void *stream);
int wm_abs_mean_bwd(const float *x, float *dx, int B, int T, float weight, int accumulate, void *stream);
/* Backward of wm_postprocess_fwd: g = dL/d delta, delta_fir = fir_lowpass(delta_raw) (wm_postprocess_fwd with
 * mode WM_POST_FIR; delta_raw itself when the FIR bit of `mode` is clear) -> d_delta_raw.  workspace: B*T floats. */
int wm_postprocess_bwd(const float *g, const float *delta_fir, const float *fir, float *d_delta_raw, void *workspace,
                       size_t workspace_bytes, int B, int T, int mode, float peak, float max_rms, float eps,
                       void *stream);

/* ---- generic fp32 operators of the main14b_2 residual stack (py/main14b_2.py:83-224, BASELINE config 3) ----
 * channels-first tensors x[b][c][t] and the reference's own parameter layouts, no packing.           */
int wm_conv1d_out_len(int Tin, int K, int stride, int pad);
int wm_convtranspose1d_out_len(int Tin, int K, int stride, int pad);
/* nn.Conv1d(Cin, Cout, K, stride, padding) (py/main14b_2.py:83-84) with fused epilogue:
 *   y = act( conv(x) + bias + chan_add[b][co] + residual ),  act 0 = identity, 1 = ELU (py/main14b_2.py:92,99-104);
 * w (Cout,Cin,K); chan_add and residual nullable.  K <= 16, stride <= 8. */
int wm_conv1d_fwd(const float *x, const float *w, const float *bias, const float *chan_add, const float *residual,
                  float *y, int B, int Cin, int Tin, int Cout, int K, int stride, int pad, int act, void *stream);
/* nn.ConvTranspose1d(Cin, Cout, K, stride, padding, output_padding=0) (py/main14b_2.py:146,201); w (Cin,Cout,K). */
int wm_convtranspose1d_fwd(const float *x, const float *w, const float *bias, float *y, int B, int Cin, int Tin,
                           int Cout, int K, int stride, int pad, void *stream);
/* The same operator for kernel_size == 2 * stride (every transposed convolution of main14b_2) as ONE stride-1
 * 3-tap convolution with Cout * stride phase channels written interleaved: `packed` (size from
 * wm_convtranspose1d_phase_weight_floats) is built once per weight version by wm_convtranspose1d_pack. */
size_t wm_convtranspose1d_phase_weight_floats(int Cin, int Cout, int stride);
int wm_convtranspose1d_pack(const float *w, const float *bias, float *packed, int Cin, int Cout, int K, int stride,
                            int pad, void *stream);
int wm_convtranspose1d_phase_fwd(const float *x, const float *packed, float *y, int B, int Cin, int Tin, int Cout, int K,
                                 int stride, int pad, void *stream);
/* nn.LSTM(H, H, num_layers, batch_first) on channels-first x[b][H][T] (py/main14b_2.py:137,167); w_ih, w_hh
 * [layers][4H][H], bias [layers][4H] = b_ih + b_hh; zero initial state; H <= 64, layers <= 4. */
int wm_lstm_small_fwd(const float *x, const float *w_ih, const float *w_hh, const float *bias, float *y, int B, int H,
                      int T, int layers, void *stream);

/* ---- the same stack on the tensor cores: planar bf16-pair activations, tcgen05 implicit GEMMs ----
 * (py/main14b_2.py:86-103 ResidualBlock incl. the strided conv1 and the 1x1 skip_conv, :146,201 ConvTranspose1d,
 * :149,205 the k7 output convolutions.)  A planar tensor of C channels (C % 8 == 0) for B clips of T steps is
 * 2 * C/8 planes of wm_pconv_plane_rows(B, T) rows x 16 bytes: plane g = bf16 hi of channels 8g..8g+7, plane
 * C/8 + g = bf16 lo (value = hi + lo); clip c, step t is row c * (T + WM_PC_GAP) + WM_PC_GAP + t, the gap rows are
 * zero (they are the convolutions' padding) and every producer below rewrites them.  The caller allocates
 * planes * plane_rows * 16 bytes with at least 128 readable bytes in front of the first plane.
 * One call computes, for every row m and output column n,
 *     acc[m][n] = bias[n] + sum over sources s, channels ci, taps j of  src_s[ci][m + row_off_s + j] * W
 * (chunk_off[n / nc] is added to the row offset of source 0), adds `residual` (planar, same geometry), applies ELU
 * when `elu`, and stores by `mode`:
 *   WM_PC_OUT_PLANAR  planar, same geometry; out_split = s > 1 writes step t to phase buffer t % s at step t / s
 *                     (buffers out_phase_rows rows apart, geometry (B, T / s); only phases 0, 1 and s - 1, the ones a
 *                     k3 stride-s convolution reads, are written);
 *   WM_PC_OUT_CONVT   column n = (phase, co) of row (c, q) is step ct_stride * q + phase of clip c in a planar tensor of
 *                     geometry (B, out_T); phase-major columns, or ct_interleave phases interleaved per 8 channels so that
 *                     a thread holds consecutive rows and stores them 32 bytes at a time;
 *   WM_PC_OUT_FP32    fp32 channels-first y[c][ch][t] for ch < ct_cout, t < out_T.
 * Weights come from wm_pconv_pack: fp32 wd[chunk][slice][16][nc] (slice = source-major, then 16-channel group, then
 * tap) -> bf16 hi | lo operand tiles. */
#define WM_PC_GAP 4
#define WM_PC_OUT_PLANAR 0
#define WM_PC_OUT_CONVT 2
#define WM_PC_OUT_FP32 3
typedef struct wm_pconv_src {
  const void *base;   /* first plane of the source (phase buffer) */
  int cin;            /* channels, multiple of 16 */
  int row_off;        /* source row of tap 0 relative to the output row */
  int taps;           /* taps on consecutive rows, 1..7 */
  int reserved;
} wm_pconv_src;
typedef struct wm_pconv {
  wm_pconv_src src[3];
  int nsrc;
  int B, T;                     /* row geometry shared by the sources, the residual and the GEMM rows */
  long long plane_rows;         /* rows per plane of the sources and the residual */
  const void *w;                /* wm_pconv_pack image */
  const float *bias;            /* [n_total] */
  int n_total, nc;              /* GEMM columns, columns per chunk (16, 32, 64, 128) */
  signed char chunk_off[64];
  int elu;
  int mode;
  const void *residual;         /* nullable */
  void *y;
  long long out_plane_rows;
  long long out_phase_rows;     /* WM_PC_OUT_PLANAR with out_split > 1: rows between phase buffers */
  int out_split;
  int ct_stride, ct_pad, ct_cout;
  int out_T;
  /* fused ResidualBlock (py/main14b_2.py:97-105), n_total == nc <= 64, WM_PC_OUT_PLANAR: the call above is conv1
   * (its result u = elu(acc) stays in shared memory), followed by
   *     y = elu( bias2 + sum over u's channels and 3 taps of u[m - 1 + j] * W2 + skip[ci][m] * Wskip + residual )
   * with w2 = wm_pconv_pack image of conv2's slices (16-channel group major, 3 taps) followed by the skip source's. */
  int fused;
  int ct_interleave;            /* WM_PC_OUT_CONVT: column n = kb * (ct_cout * k) + g * (8 k) + phl * 8 + c is phase kb * k + phl of
                                 * channel 8 g + c (k = ct_interleave, a divisor of ct_stride; 0 or 1: n = phase * ct_cout + co) */
  const void *w2;
  const float *bias2;
  wm_pconv_src skip;            /* base == NULL: none (one tap, row offset 0: phase 0 of the block input) */
} wm_pconv;
long long wm_pconv_plane_rows(int B, int T);
size_t wm_pconv_desc_bytes(void);   /* sizeof(wm_pconv), for bindings that mirror the struct */
size_t wm_pconv_weight_bytes(long long nslices, int nc);
int wm_pconv_pack(const float *wd, void *img, long long nslices, int nc, void *stream);
int wm_pconv_fwd(const wm_pconv *desc, void *stream);
/* Conv1d(1, cout, K, padding K/2) on waveforms s[B][T] (py/main14b_2.py:123,190) -> planar, split into `split` phases */
int wm_pconv_in_fwd(const float *s, const float *w, const float *bias, void *y, int B, int T, int cout, int K, int split,
                    long long plane_rows, void *stream);
/* The Generator's last ResidualBlock(8, 8) + final_conv_dec Conv1d(8, 1, 7, padding 3) + crop to T
 * (py/main14b_2.py:97-105,149,173-177) in one kernel: x planar (8 channels, geometry (B, Tx)), reference weight layouts
 * w1, w2 (8,8,3), wf (1,8,7) -> delta[B][T] (zero beyond Tx). */
int wm_m14_tail8_fwd(const void *x, long long plane_rows, int B, int Tx, const float *w1, const float *b1, const float *w2,
                     const float *b2, const float *wf, const float *bf, float *delta, int T, void *stream);
/* fp32 channels-first x[B][C][T] <-> planar */
int wm_pconv_to_planar(const float *x, void *y, int B, int C, int T, long long plane_rows, void *stream);
int wm_pconv_from_planar(const void *x, float *y, int B, int C, int T, int Tout, long long plane_rows, void *stream);

/* ---- audio formats either side of the path ----
 * torchaudio.transforms.Resample(orig, new)(x) (py/main16.py:985,1121): `kern` [K][up] is torchaudio's windowed-sinc
 * table transposed (packing.resample_kernel builds it as torchaudio does), down/up = orig/new divided by their gcd,
 * y[b][m*up + j] = sum_k kern[k][j] x[b][m*down + k - width]; Tout = ceil(up * Tin / down). */
int wm_resample_fwd(const float *x, const float *kern, float *y, int B, int Tin, int Tout, int down, int up, int K,
                    int width, void *stream);
/* (clamp(x,-1,1) * 32767).to(int16) (py/main15.py:859-860) and the PCM loader's int16 * scale. */
int wm_pcm16_quantize_fwd(const float *x, int16_t *q, size_t n, void *stream);
int wm_pcm16_dequantize_fwd(const int16_t *q, float *x, size_t n, float scale, void *stream);
/* Per-row quality metrics of generate_watermarked_audio (py/main16.py:1030-1049; compute_si_snr :764-773):
 * out[b] = {watermark_rms, si_snr_db, power_ratio_db} over the first valid_len[b] samples (nullable = T). */
int wm_file_metrics_fwd(const float *s, const float *s_w, const int *valid_len, float *out, int B, int T, void *stream);

/* Second-order IIR with torchaudio.functional.lfilter semantics (py/main15.py:855 lowpass_biquad; main15c.ipynb
 * cell 4): y[n] = (b0 x[n] + b1 x[n-1] + b2 x[n-2] - a1 y[n-1] - a2 y[n-2]) / a0 per row of x[rows][N], zero initial
 * state, clamp != 0 clamps y to [-1, 1] (lfilter's default); pcm16 (optional) additionally receives
 * (int16)(clamp(y, -1, 1) * 32767) (py/main15.py:859-860).  b3 / a3: HOST pointers to 3 doubles each.  The
 * recurrence is evaluated as a chunked scan in double precision.  Not in place. */
size_t wm_biquad_workspace_bytes(int rows, long long N);
int wm_biquad_fwd(const float *x, float *y, int16_t *pcm16, int rows, long long N, const double *b3, const double *a3,
                  int clamp, void *workspace, size_t workspace_bytes, void *stream);

/* Detection statistics of the evaluation cells, on device scores.
 * confusion_counts: out4 = {tn, fp, fn, tp} with prediction = score >= thresh (py/main16.py:1335-1341).
 * roc_points: for every threshold t: fp[t] = #{clean >= t}, tp[t] = #{wm >= t} (the points sklearn's roc_curve
 *             returns for these thresholds, py/main16.py:2372-2386).
 * auc_pairs:  out = 2 * #{wm > clean} + #{wm == clean} over all pairs; AUC = out / (2 n_clean n_wm). */
int wm_confusion_counts_fwd(const float *clean, long long n_clean, const float *wm, long long n_wm, float thresh,
                            unsigned long long *out4, void *stream);
int wm_roc_points_fwd(const float *clean, long long n_clean, const float *wm, long long n_wm, const float *thresholds,
                      int n_thresholds, int *fp, int *tp, void *stream);
int wm_auc_pairs_fwd(const float *clean, long long n_clean, const float *wm, long long n_wm, unsigned long long *out,
                     void *stream);

/* ---- training (BASELINE config 4): the Detector's half of train_one_epoch ----
 * One optimisation step of the Detector (py/main16.py:160-176 in train mode: batch-statistics BatchNorm, running
 * stats updated with momentum 0.1) on x[B_total][T] whose first B_wm clips are watermarked with message[b]
 * (py/main16.py:249-264): loss = lam_loc * BCE(ch 0, [b < B_wm]) + lam_dec * BCE(ch 1.., bits), backward through
 * the whole network, then torch.optim.Adam (bias-corrected, no weight decay) when adam_step >= 1 (adam_step = 0
 * leaves params untouched: gradients only).  params / grads / adam_m / adam_v: WM_DT_SIZE floats, run_stats:
 * WM_DT_STATS floats, losses_out (device, nullable): {loc, bce}; d_input (nullable) [B_total][T] receives the
 * gradient w.r.t. x (what the generator's backward consumes).  fp32 on the CUDA cores, deterministic. */
size_t wm_detector_train_workspace_bytes(int B_total, int T, int nout);
int wm_detector_train_step(float *params, float *grads, float *adam_m, float *adam_v, float *run_stats, const float *x,
                           const int64_t *message, int B_wm, int B_total, int T, int nout, float lam_loc, float lam_dec,
                           float lr, float beta1, float beta2, float eps, int adam_step, float *losses_out,
                           float *d_input, void *workspace, size_t workspace_bytes, void *stream);
/* Forward + backward of one train_one_epoch iteration (py/main16.py:244-277), both networks in train mode:
 * delta = limit_rms(clamp_peak(fir_lowpass(G(s, message)))), s_w = s + delta, logits = D(cat(s_w, s)),
 * loss = lam[0] l1 + lam[1] mel + lam[2] loud + lam[3] loc + lam[4] bce + lam[5] hf  (lam on the HOST).
 * Gradients land in g_grads (WM_GT_SIZE, dense embedding rows included) and d_grads (WM_DT_SIZE); the BatchNorm
 * running stats (g_stats WM_GT_STATS, d_stats WM_DT_STATS) are updated; parameters are NOT touched — the caller
 * all-reduces the gradient buffers across ranks if it wants to and then calls wm_adam_step on each
 * (py/main16.py:278,504).  losses_out[8] (device): {l1, mel, loud, loc, bce, hf, total, raw_total};
 * s_w_out (nullable) [B][T].  mel_fb / mel_band as for wm_mel_log_l1_fwd (n_fft 1024). */
size_t wm_train_step_workspace_bytes(int B, int T, int nout);
int wm_train_forward_backward(const float *g_params, float *g_grads, float *g_stats, const float *d_params,
                              float *d_grads, float *d_stats, const float *s, const int64_t *message, const float *fir,
                              const float *mel_fb, const int *mel_band, int n_mels, const float *lam, int B, int T,
                              int nout, float *losses_out, float *s_w_out, void *workspace, size_t workspace_bytes,
                              void *stream);
/* Building blocks of the step, exposed for the parity tests (channels-last x[n][64], n = B*T rows):
 * nn.BatchNorm1d(64) in train mode fused with the residual add and ReLU of py/main16.py:118-127 ... */
int wm_bn_train_fwd(const float *z, const float *gamma, const float *beta, const float *residual, float *out,
                    float *mean, float *rstd, float *run_mean, float *run_var, long long rows, int relu,
                    void *workspace, size_t workspace_bytes, void *stream);
/* ... its backward: dout is the gradient of `act` (the block output after ReLU, nullable = no ReLU); dres (nullable)
 * receives the gradient of the residual input. */
int wm_bn_train_bwd(const float *dout, const float *act, const float *z, const float *mean, const float *rstd,
                    const float *gamma, float *dz, float *dres, float *dgamma, float *dbeta, long long rows,
                    void *workspace, size_t workspace_bytes, void *stream);
/* Gradients of y = Conv1d(64,64,K,padding=K/2)(x) given dy: dw [K][64][64] tap-major, db [64], dx (nullable; needs
 * w, the forward weight in the same layout).  K in {1,3,7}. */
size_t wm_conv64_bwd_workspace_bytes(int B, int T, int K);
/* The training step's forward convolution as a single operator: y = conv(x) + bias (+ residual), fp32 channels-last in
 * and out.  In the default math mode (WM_MATH_BF16X2) forward, data gradient and weight gradient of these
 * convolutions run on tcgen05 (bf16 hi+lo pairs, fp32 accumulate); WM_MATH_FP32 selects the fp32 FMA kernels.
 * Workspace: wm_conv64_bwd_workspace_bytes(B, T, K) for both. */
int wm_conv64_train_fwd(const float *x, const float *w, const float *bias, const float *residual, float *y, int B, int T,
                        int K, void *workspace, size_t workspace_bytes, void *stream);
int wm_conv64_bwd(const float *x, const float *dy, const float *w, float *dw, float *db, float *dx, int B, int T, int K,
                  void *workspace, size_t workspace_bytes, void *stream);
/* Backward of the remaining single operators (each checked against autograd in tests/test_train.py):
 * wm_bce_heads_bwd   d(lam_loc * loc + lam_dec * bce)/d logits for logits[B_total][T][nout] (py/main16.py:255-264);
 * wm_head_bwd        Conv1d(64,nout,1) on channels-last y[rows][64]: dy, dw [nout][64], db [nout] from dlogits;
 * wm_conv_in_k7_bwd  Conv1d(1,64,7,padding=3): dw [7][64], db [64] and (nullable) ds [B][T] from dx [B][T][64]. */
int wm_bce_heads_bwd(const float *logits, const int64_t *message, float *dlogits, int B_wm, int B_total, int T, int nout,
                     float lam_loc, float lam_dec, void *stream);
size_t wm_head_bwd_workspace_bytes(long long rows, int nout);
int wm_head_bwd(const float *dlogits, const float *y, const float *w, float *dy, float *dw, float *db, long long rows,
                int nout, void *workspace, size_t workspace_bytes, void *stream);
size_t wm_conv_in_k7_bwd_workspace_bytes(int B, int T);
int wm_conv_in_k7_bwd(const float *s, const float *dx, const float *w, float *dw, float *db, float *ds, int B, int T,
                      void *workspace, size_t workspace_bytes, void *stream);
size_t wm_bn_train_workspace_bytes(long long rows);
/* nn.LSTM(64,64,batch_first) in training (py/main16.py:138,153).  Weights "per-gate transposed":
 * wT[q][k][r] = W[q*64 + r][k], q over (i,f,g,o).  Forward keeps the activated gates [B][T][256] and cell states
 * [B][T][64]; backward turns dy into dx, dwT_ih, dwT_hh [4][64][64] and db [256] (gradient of b_ih and of b_hh). */
int wm_lstm_train_fwd(const float *x, const float *wT_ih, const float *wT_hh, const float *b_ih, const float *b_hh,
                      float *h, float *gates, float *cell, int B, int T, void *stream);
size_t wm_lstm_train_bwd_workspace_bytes(int B, int T);
int wm_lstm_train_bwd(const float *dy, const float *x, const float *h, const float *wT_ih, const float *wT_hh,
                      const float *gates, const float *cell, float *dx, float *dwT_ih, float *dwT_hh, float *db, int B,
                      int T, void *workspace, size_t workspace_bytes, void *stream);
/* torch.optim.Adam step `step` (>= 1) on n floats. */
int wm_adam_step(float *p, const float *g, float *m, float *v, long long n, float lr, float beta1, float beta2,
                 float eps, int step, void *stream);

/* Same unit with HOST (pinned) buffers: H2D of s and message, the device pipeline in
 * micro-batches of `chunk` clips, D2H of s_w, probs, clip_prob and msg_logits.  This is what
 * the file-level API and bench.py's e2e leg call.  The device staging area lives in the
 * caller's workspace (wm_embed_detect_host_workspace_bytes).
 * Stream contract: the kernels run on `stream`; the copies run on two copy streams (in / out)
 * that the LIBRARY owns per host thread (created on first use, never destroyed) so that they
 * overlap the kernels — a deviation from "the caller owns everything", made because the
 * overlap needs three streams and the reference-facing call has one.  The call orders them
 * with events: the copy-in stream waits for everything queued on `stream` before the call,
 * and before returning `stream` is made to wait for every copy of this call, so work the
 * caller queues on `stream` afterwards sees the results and may reuse the buffers.  The call
 * itself does not block the host. */
size_t wm_embed_detect_host_workspace_bytes(int chunk, int T, int nout);
int wm_embed_detect_host(const float *g_blob, const float *embedding, int64_t emb_rows,
                         const float *d_blob, const float *fir,
                         const int64_t *host_message, const float *host_s,
                         float *host_s_w, float *host_probs, float *host_clip_prob,
                         float *host_msg_logits, void *workspace, size_t workspace_bytes,
                         int B, int T, int nout, int chunk, int post_mode, void *stream);
/* The same call on a caller's buffer that holds only `host_s_floats` <= B * T samples: the missing tail of the last
 * clip(s) is zero on the device (py/main16.py:1011-1026 right-zero-pads the last partial segment).  Lets the long-form
 * drivers run a ragged recording straight out of its own pinned memory, with no staging copy. */
int wm_embed_detect_host_ragged(const float *g_blob, const float *embedding, int64_t emb_rows,
                                const float *d_blob, const float *fir,
                                const int64_t *host_message, const float *host_s, long long host_s_floats,
                                float *host_s_w, float *host_probs, float *host_clip_prob,
                                float *host_msg_logits, void *workspace, size_t workspace_bytes,
                                int B, int T, int nout, int chunk, int post_mode, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* WMB200_H */
