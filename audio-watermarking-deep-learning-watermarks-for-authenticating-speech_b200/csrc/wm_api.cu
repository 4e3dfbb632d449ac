// C ABI of libwmb200 (include/wmb200.h): argument checking, error reporting and the
// module-level drivers that string the kernels together on the caller's stream.
#include <stdarg.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "wm_common.h"

namespace wm {

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};
static std::atomic<int> g_math_mode{WM_MATH_BF16X2};

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int math_mode() { return g_math_mode.load(); }
void count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

int sm_count() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

int require_device() {
  static int ok = -1;
  if (ok < 0) {
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
      set_error("no CUDA device: %s", cudaGetErrorString(cudaGetLastError()));
      return -4;
    }
    ok = (major == 10) ? 1 : 0;
  }
  if (!ok) {
    set_error("libwmb200 is built for sm_100a only and has no fallback path");
    return -4;
  }
  return 0;
}

// Host copy of the fp32 parameter block of every finalized blob, keyed by the blob's device address.  The fused
// kernels take biases and 1x1-head weights as a kernel parameter (constant bank), which needs them on the host;
// wm_finalize_*_blob reads the block back once (that one-time set-up call synchronises its stream; nothing on the
// per-batch path does).  A blob must not be modified after it has been finalized.
struct HostBlob {
  std::vector<float> v;                       // the fp32 block (WM_G_SIZE or WM_D_SIZE floats)
  float head0[65];                            // output 0 of the 1x1 head: w[64] then b
  float head_all[WM_MAX_HEAD * 65];           // detector: w[nout_max][64] then b[nout_max] is rebuilt per nout on use
  const float *at(int off) const { return v.data() + off; }
};
static std::mutex g_head_mu;
static std::unordered_map<const float *, HostBlob> g_heads;
static int remember_blob(const float *blob, int nfloats, int w_off, int b_off, cudaStream_t st) {
  HostBlob h;
  h.v.resize(nfloats);
  WM_CHECK_CUDA(cudaMemcpyAsync(h.v.data(), blob, sizeof(float) * nfloats, cudaMemcpyDeviceToHost, st));
  WM_CHECK_CUDA(cudaStreamSynchronize(st));
  memcpy(h.head0, h.at(w_off), sizeof(float) * 64);
  h.head0[64] = *h.at(b_off);
  std::lock_guard<std::mutex> lk(g_head_mu);
  g_heads[blob] = std::move(h);
  return 0;
}
// the map never erases: the returned pointer stays valid for the life of the process
static const HostBlob *lookup_blob(const float *blob) {
  std::lock_guard<std::mutex> lk(g_head_mu);
  auto it = g_heads.find(blob);
  return it == g_heads.end() ? nullptr : &it->second;
}
// b1[64] then b2[64] of the ResBlock at float offset `rb` (host), for the by-value kernel parameter
static void host_bias_pair(const HostBlob *h, int rb, float *out128) {
  memcpy(out128, h->at(rb + WM_RB_B1), sizeof(float) * 64);
  memcpy(out128 + 64, h->at(rb + WM_RB_B2), sizeof(float) * 64);
}

static size_t align256(size_t n) { return (n + 255) & ~(size_t)255; }
static size_t planar_bytes(int B, int T) {
  // + 4 KB: the bulk copies of a partial last tile read (never use) up to 127 rows past a plane
  return align256((size_t)B * 16 * ((size_t)T + 2 * WM_PLANAR_PAD) * 16 + 4096);
}
// one activation buffer: large enough for either layout (fp32 channels-last or planar bf16 pairs)
static size_t act_bytes(int B, int T) { return planar_bytes(B, T); }

struct Ws {
  char *p;
  size_t left;
  void *take(size_t n) {
    n = align256(n);
    if (n > left) return nullptr;
    void *r = p;
    p += n;
    left -= n;
    return r;
  }
};

// ResBlock (py/main16.py:112-125): y = relu(x + conv2(relu(conv1(x)))) with BN folded, fp32 FMA.
static int resblock_fp32(const float *rb, const float *x, float *tmp, float *y, int B, int T, cudaStream_t st) {
  WM_TRY(launch_conv64_fp32(x, rb + WM_RB_W1, rb + WM_RB_B1, nullptr, nullptr, tmp, B, T, 3, 1, st));
  WM_TRY(launch_conv64_fp32(tmp, rb + WM_RB_W2, rb + WM_RB_B2, x, nullptr, y, B, T, 3, 1, st));
  return 0;
}
// the same on planar tensors with the tcgen05 kernel; the second conv writes planar `y` and/or fp32 `y32`
static int resblock_tc(const float *blob, int rb_off, const float *img /* two 3-tap images */, const void *x, void *tmp,
                       void *y, float *y32, int B, int T, cudaStream_t st) {
  (void)tmp;
  const float *rb = blob + rb_off;
  float hb[128];
  const HostBlob *h = lookup_blob(blob);
  if (h) host_bias_pair(h, rb_off, hb);
  return launch_resblock_tc(x, img, rb + WM_RB_B1, rb + WM_RB_B2, y, y32, B, T, st, h ? hb : nullptr);
}

// The tcgen05 generator in three phases, so that a host-fed pipeline can run the two convolutional phases
// per sub-batch (overlapping their neighbours' PCIe copies) around one whole-batch LSTM launch.
// encoder: conv k7 -> ResBlock -> ResBlock (py/main16.py:133-137); s[B][T] -> r1 planar (r0, r2 scratch)
static int generator_encoder_tc(const float *blob, const float *s, void *r0, void *r1, void *r2, int B, int T,
                                cudaStream_t st) {
  const float *tc = blob + WM_G_TC;
  (void)r0;   // input convolution folded into the first ResBlock: s -> r2 planar
  WM_TRY(launch_resblock_in_tc(s, blob + WM_G_FIN + WM_FIN_W9, blob + WM_G_IN_W, blob + WM_G_FIN, tc + WM_TC_IMG3,
                               blob + WM_G_RB0 + WM_RB_B2, r2, B, T, st));
  return resblock_tc(blob, WM_G_RB1, tc + 2 * WM_TC_IMG3, r2, r0, r1, nullptr, B, T, st);      // -> r1 planar
}
// LSTM (+ message embedding added to its output)            (py/main16.py:152-159); x planar -> y planar
static int generator_lstm_tc(const float *blob, const float *chan_add, const void *x, void *y, int B, int T,
                             cudaStream_t st) {
  return launch_lstm_tc(x, blob + WM_G_TC_LSTM_W, blob + WM_G_TC_LSTM_B, chan_add, y, B, T, st);
}
// decoder: ConvTranspose k7 -> ResBlock with the Conv1d(64,1,1) head fused into its epilogue (py/main16.py:143-147)
static int generator_decoder_tc(const float *blob, const void *x, void *tmp, float *delta_raw, int B, int T,
                                cudaStream_t st) {
  WM_TRY(launch_conv64_tc(x, blob + WM_G_TC_CT, blob + WM_G_CT_B, nullptr, tmp, nullptr, B, T, 7, 0, st));
  if (const HostBlob *h = lookup_blob(blob)) {
    float hb[128];
    host_bias_pair(h, WM_G_RB2, hb);
    return launch_resblock_head1_tc(tmp, blob + WM_G_TC_RB2, blob + WM_G_RB2 + WM_RB_B1, blob + WM_G_RB2 + WM_RB_B2,
                                    h->head0, delta_raw, B, T, st, hb);
  }
  // blob never went through wm_finalize_generator_blob on this process: un-fused head (x is dead by now)
  float *f = (float *)const_cast<void *>(x);
  WM_TRY(launch_resblock_tc(tmp, blob + WM_G_TC_RB2, blob + WM_G_RB2 + WM_RB_B1, blob + WM_G_RB2 + WM_RB_B2, nullptr, f, B,
                            T, st));
  return launch_head(f, blob + WM_G_HEAD_W, blob + WM_G_HEAD_B, delta_raw, B, T, 1, st);
}

// r0, r1, r2: three activation buffers of act_bytes(B, T); result: delta_raw[B][T]
static int generator_run(const float *blob, const float *embedding, int64_t emb_rows,
                         const int64_t *message, const float *s, float *delta_raw, void *r0, void *r1,
                         void *r2, float *emb, int B, int T, cudaStream_t st) {
  const float *chan_add = nullptr;
  if (message && embedding) {  // embedding(message)  (py/main16.py:156-158)
    WM_TRY(launch_gather_rows(embedding, emb_rows, message, emb, B, st));
    chan_add = emb;
  }
  float *f0 = (float *)r0, *f1 = (float *)r1, *f2 = (float *)r2;
  if (g_math_mode.load() == WM_MATH_FP32) {
    // encoder: conv k7 -> ResBlock -> ResBlock                 (py/main16.py:133-137)
    WM_TRY(launch_conv_in_k7(s, blob + WM_G_IN_W, blob + WM_G_IN_B, f0, B, T, st));
    WM_TRY(resblock_fp32(blob + WM_G_RB0, f0, f1, f2, B, T, st));  // -> r2
    WM_TRY(resblock_fp32(blob + WM_G_RB1, f2, f0, f1, B, T, st));  // -> r1
    // LSTM over time                                             (py/main16.py:152-154)
    WM_TRY(launch_lstm_fp32(f1, blob + WM_G_LSTM_WIH, blob + WM_G_LSTM_WHH, blob + WM_G_LSTM_B, f0, B, T, st));
    // decoder: (+ message) ConvTranspose k7 -> ResBlock -> Conv 64->1   (py/main16.py:143-147,156-159)
    WM_TRY(launch_conv64_fp32(f0, blob + WM_G_CT_W, blob + WM_G_CT_B, nullptr, chan_add, f2, B, T, 7, 0, st));
    WM_TRY(resblock_fp32(blob + WM_G_RB2, f2, f0, f1, B, T, st));  // -> r1
    return launch_head(f1, blob + WM_G_HEAD_W, blob + WM_G_HEAD_B, delta_raw, B, T, 1, st);
  }
  WM_TRY(generator_encoder_tc(blob, s, r0, r1, r2, B, T, st));
  WM_TRY(generator_lstm_tc(blob, chan_add, r1, r0, B, T, st));
  return generator_decoder_tc(blob, r0, r1, delta_raw, B, T, st);
}

// Detector trunk (py/main16.py:176-179): fp32 channels-last result left in *out (one of the buffers)
static int detector_trunk(const float *blob, const float *x, void *r0, void *r1, void *r2,
                          float **out, int B, int T, cudaStream_t st) {
  float *f0 = (float *)r0, *f1 = (float *)r1, *f2 = (float *)r2;
  *out = f1;
  if (g_math_mode.load() == WM_MATH_FP32) {
    WM_TRY(launch_conv_in_k7(x, blob + WM_D_IN_W, blob + WM_D_IN_B, f0, B, T, st));
    WM_TRY(resblock_fp32(blob + WM_D_RB0, f0, f1, f2, B, T, st));  // -> r2
    return resblock_fp32(blob + WM_D_RB1, f2, f0, f1, B, T, st);   // -> r1
  }
  const float *tc = blob + WM_D_TC;
  (void)r0;   // input convolution folded into the first ResBlock: x -> r2 planar
  WM_TRY(launch_resblock_in_tc(x, blob + WM_D_FIN + WM_FIN_W9, blob + WM_D_IN_W, blob + WM_D_FIN, tc + WM_TC_IMG3,
                               blob + WM_D_RB0 + WM_RB_B2, r2, B, T, st));
  return resblock_tc(blob, WM_D_RB1, tc + 2 * WM_TC_IMG3, r2, r0, nullptr, f1, B, T, st);
}

// Detector + heads (py/main16.py:176-180,1142-1146): per-sample probability, clip mean, mean message logits.
// tcgen05 mode without a vote request: channel 0 of the 1x1 head, the sigmoid and per-tile partial sums of the
// probability and of the 64 activations run in the last ResBlock's epilogue (the (B,T,64) feature map is never
// stored); the message head is applied to the per-clip activation means by the finalize kernel.
static int detect_run(const float *blob, const float *x, const int *valid_len, float *probs, float *clip_prob,
                      float *msg_logits, float *vote_frac, void *r0, void *r1, void *r2, int B, int T, int nout,
                      cudaStream_t st) {
  const HostBlob *h = lookup_blob(blob);
  if (g_math_mode.load() == WM_MATH_BF16X2 && (vote_frac == nullptr || nout <= WM_FUSED_VOTE_MAX) &&
      (size_t)B * resblock_tiles_per_clip(T) * 4 * WM_DET_PART * sizeof(float) <= act_bytes(B, T) && h != nullptr) {
    const float *tc = blob + WM_D_TC;
    float *partials = (float *)r1;
    (void)r0;   // input convolution folded into the first ResBlock: x -> r2 planar
    WM_TRY(launch_resblock_in_tc(x, blob + WM_D_FIN + WM_FIN_W9, blob + WM_D_IN_W, blob + WM_D_FIN, tc + WM_TC_IMG3,
                                 blob + WM_D_RB0 + WM_RB_B2, r2, B, T, st));
    float hb[128];
    host_bias_pair(h, WM_D_RB1, hb);
    const bool votes = vote_frac != nullptr && nout > 1;
    float head_all[WM_MAX_HEAD * 65];
    if (votes) {   // rows 0..nout-1 of the head, then their biases (majority vote in the epilogue)
      memcpy(head_all, h->at(WM_D_HEAD_W), sizeof(float) * 64 * nout);
      memcpy(head_all + 64 * nout, h->at(WM_D_HEAD_B), sizeof(float) * nout);
    }
    WM_TRY(launch_resblock_detect_tc(r2, tc + 2 * WM_TC_IMG3, blob + WM_D_RB1 + WM_RB_B1, blob + WM_D_RB1 + WM_RB_B2,
                                     h->head0, valid_len, probs, partials, B, T, st, hb, votes ? head_all : nullptr,
                                     votes ? nout : 0));
    return launch_detect_finalize(partials, valid_len, blob + WM_D_HEAD_W, blob + WM_D_HEAD_B, clip_prob, msg_logits,
                                  votes ? vote_frac : nullptr, B, T, nout, st);
  }
  float *out = nullptr;
  WM_TRY(detector_trunk(blob, x, r0, r1, r2, &out, B, T, st));
  return launch_head_detect(out, blob + WM_D_HEAD_W, blob + WM_D_HEAD_B, valid_len, probs, clip_prob, msg_logits,
                            vote_frac, B, T, nout, st);
}

}  // namespace wm

using namespace wm;

extern "C" {

int wm_abi_version(void) { return WM_ABI_VERSION; }
const char *wm_last_error(void) { return g_err; }
int wm_device_ok(void) { return require_device() == 0 ? 1 : 0; }
int wm_set_math_mode(int mode) {
  int prev = g_math_mode.load();
  if (mode == WM_MATH_FP32 || mode == WM_MATH_BF16X2) g_math_mode.store(mode);
  return prev;
}
int wm_get_math_mode(void) { return g_math_mode.load(); }
unsigned long long wm_launch_count(void) { return g_launches.load(); }

#define WM_ENTRY()                 \
  do {                             \
    int rc0_ = require_device();   \
    if (rc0_ != 0) return rc0_;    \
  } while (0)

int wm_finalize_generator_blob(float *blob, void *stream) {
  WM_ENTRY();
  WM_CHECK_ARG(blob, "finalize_generator_blob: null pointer");
  cudaStream_t st = as_stream(stream);
  const int rb[3] = {WM_G_RB0, WM_G_RB1, WM_G_RB2};
  const int img[3] = {WM_G_TC, WM_G_TC + 2 * WM_TC_IMG3, WM_G_TC_RB2};
  for (int i = 0; i < 3; ++i) {
    WM_TRY(launch_pack_conv64_tc(blob + rb[i] + WM_RB_W1, blob + img[i], 3, st));
    WM_TRY(launch_pack_conv64_tc(blob + rb[i] + WM_RB_W2, blob + img[i] + WM_TC_IMG3, 3, st));
  }
  WM_TRY(launch_pack_conv64_tc(blob + WM_G_CT_W, blob + WM_G_TC_CT, 7, st));
  WM_TRY(launch_pack_lstm_tc(blob + WM_G_LSTM_WIH, blob + WM_G_LSTM_WHH, blob + WM_G_LSTM_B, blob + WM_G_TC_LSTM_W,
                             blob + WM_G_TC_LSTM_B, st));
  return remember_blob(blob, WM_G_SIZE, WM_G_HEAD_W, WM_G_HEAD_B, st);   // synchronises `stream` (one-time set-up call)
}

int wm_finalize_detector_blob(float *blob, void *stream) {
  WM_ENTRY();
  WM_CHECK_ARG(blob, "finalize_detector_blob: null pointer");
  cudaStream_t st = as_stream(stream);
  const int rb[2] = {WM_D_RB0, WM_D_RB1};
  for (int i = 0; i < 2; ++i) {
    WM_TRY(launch_pack_conv64_tc(blob + rb[i] + WM_RB_W1, blob + WM_D_TC + (2 * i) * WM_TC_IMG3, 3, st));
    WM_TRY(launch_pack_conv64_tc(blob + rb[i] + WM_RB_W2, blob + WM_D_TC + (2 * i + 1) * WM_TC_IMG3, 3, st));
  }
  return remember_blob(blob, WM_D_SIZE, WM_D_HEAD_W, WM_D_HEAD_B, st);   // synchronises `stream` (one-time set-up call)
}

size_t wm_planar_bytes(int B, int T) { return (B <= 0 || T <= 0) ? 0 : planar_bytes(B, T); }
size_t wm_conv64_tc_weight_bytes(int taps) { return (size_t)taps * 8 * 128 * 8 * 2; }

int wm_to_planar(const float *x, const float *chan_add, void *y, int B, int T, void *stream) {
  WM_ENTRY();
  WM_CHECK_ARG(B >= 0 && T >= 0, "to_planar: negative size");
  WM_CHECK_ARG(B == 0 || T == 0 || (x && y), "to_planar: null pointer");
  return launch_to_planar(x, chan_add, y, B, T, as_stream(stream));
}

int wm_from_planar(const void *x, float *y, int B, int T, void *stream) {
  WM_ENTRY();
  WM_CHECK_ARG(B >= 0 && T >= 0, "from_planar: negative size");
  WM_CHECK_ARG(B == 0 || T == 0 || (x && y), "from_planar: null pointer");
  return launch_from_planar(x, y, B, T, as_stream(stream));
}

int wm_pack_conv64_tc(const float *w, void *img, int taps, void *stream) {
  WM_ENTRY();
  WM_CHECK_ARG(w && img, "pack_conv64_tc: null pointer");
  WM_CHECK_ARG(taps == 3 || taps == 7, "pack_conv64_tc: taps must be 3 or 7");
  return launch_pack_conv64_tc(w, img, taps, as_stream(stream));
}

int wm_conv64_tc_fwd(const void *x, const void *w_img, const float *bias, const void *residual, void *y,
                     float *y32, int B, int T, int taps, int relu, void *stream) {
  WM_ENTRY();
  WM_CHECK_ARG(B >= 0 && T >= 0, "conv64_tc: negative size");
  WM_CHECK_ARG(B == 0 || T == 0 || (x && w_img && bias && (y || y32)), "conv64_tc: null pointer");
  WM_CHECK_ARG(x != y, "conv64_tc: in-place operation is not supported");
  return launch_conv64_tc(x, w_img, bias, residual, y, y32, B, T, taps, relu, as_stream(stream));
}

/* developer hook (not part of the ABI contract): per-phase cycle counters of the LSTM kernel */
int wm_debug_lstm_profile(long long *buf16) {
  set_lstm_profile_buffer(buf16);
  return 0;
}
int wm_debug_lstm_opts(int opts) {
  set_lstm_opts(opts);
  return 0;
}

int wm_resblock_tc_fwd(const void *x, const void *w_img, const float *b1, const float *b2, void *y, float *y32,
                       int B, int T, void *stream) {
  WM_ENTRY();
  WM_CHECK_ARG(B >= 0 && T >= 0, "resblock_tc: negative size");
  WM_CHECK_ARG(B == 0 || T == 0 || (x && w_img && b1 && b2 && (y || y32)), "resblock_tc: null pointer");
  WM_CHECK_ARG(x != y, "resblock_tc: in-place operation is not supported");
  return launch_resblock_tc(x, w_img, b1, b2, y, y32, B, T, as_stream(stream));
}

int wm_resblock_tc_hostbias_fwd(const void *x, const void *w_img, const float *host_b1, const float *host_b2, void *y,
                                float *y32, int B, int T, void *stream) {
  WM_ENTRY();
  WM_CHECK_ARG(B >= 0 && T >= 0, "resblock_tc_hostbias: negative size");
  WM_CHECK_ARG(B == 0 || T == 0 || (x && w_img && host_b1 && host_b2 && (y || y32)), "resblock_tc_hostbias: null pointer");
  WM_CHECK_ARG(x != y, "resblock_tc_hostbias: in-place operation is not supported");
  float hb[128];
  memcpy(hb, host_b1, sizeof(float) * 64);
  memcpy(hb + 64, host_b2, sizeof(float) * 64);
  return launch_resblock_tc(x, w_img, nullptr, nullptr, y, y32, B, T, as_stream(stream), hb);
}

int wm_pack_lstm_tc(const float *w_ih, const float *w_hh, const float *bias, void *wpk, float *bias_p,
                    void *stream) {
  WM_ENTRY();
  WM_CHECK_ARG(w_ih && w_hh && bias && wpk && bias_p, "pack_lstm_tc: null pointer");
  return launch_pack_lstm_tc(w_ih, w_hh, bias, wpk, bias_p, as_stream(stream));
}

int wm_lstm_tc_fwd(const void *x, const void *wpk, const float *bias_p, const float *chan_add, void *y, int B,
                   int T, void *stream) {
  WM_ENTRY();
  WM_CHECK_ARG(B >= 0 && T >= 0, "lstm_tc: negative size");
  WM_CHECK_ARG(B == 0 || T == 0 || (x && wpk && bias_p && y), "lstm_tc: null pointer");
  WM_CHECK_ARG(x != y, "lstm_tc: in-place operation is not supported");
  return launch_lstm_tc(x, wpk, bias_p, chan_add, y, B, T, as_stream(stream));
}

int wm_conv_in_k7_fwd(const float *s, const float *w, const float *b, float *y, int B, int T,
                      void *stream) {
  WM_ENTRY();
  WM_CHECK_ARG(B >= 0 && T >= 0, "conv_in_k7: negative size");
  WM_CHECK_ARG(B == 0 || T == 0 || (s && w && b && y), "conv_in_k7: null pointer");
  return launch_conv_in_k7(s, w, b, y, B, T, as_stream(stream));
}

int wm_conv64_fwd(const float *x, const float *w, const float *bias, const float *residual,
                  const float *chan_add, float *y, int B, int T, int taps, int relu, void *stream) {
  WM_ENTRY();
  WM_CHECK_ARG(B >= 0 && T >= 0, "conv64: negative size");
  WM_CHECK_ARG(B == 0 || T == 0 || (x && w && bias && y), "conv64: null pointer");
  WM_CHECK_ARG(x != y, "conv64: in-place operation is not supported");
  return launch_conv64_fp32(x, w, bias, residual, chan_add, y, B, T, taps, relu, as_stream(stream));
}

int wm_lstm_fwd(const float *x, const float *w_ih, const float *w_hh, const float *bias, float *h,
                int B, int T, void *stream) {
  WM_ENTRY();
  WM_CHECK_ARG(B >= 0 && T >= 0, "lstm: negative size");
  WM_CHECK_ARG(B == 0 || T == 0 || (x && w_ih && w_hh && bias && h), "lstm: null pointer");
  WM_CHECK_ARG(x != h, "lstm: in-place operation is not supported");
  return launch_lstm_fp32(x, w_ih, w_hh, bias, h, B, T, as_stream(stream));
}

int wm_head_fwd(const float *x, const float *w, const float *b, float *y, int B, int T, int nout,
                void *stream) {
  WM_ENTRY();
  WM_CHECK_ARG(B >= 0 && T >= 0, "head: negative size");
  WM_CHECK_ARG(B == 0 || T == 0 || (x && w && b && y), "head: null pointer");
  return launch_head(x, w, b, y, B, T, nout, as_stream(stream));
}

int wm_postprocess_fwd(const float *delta_raw, const float *s, const float *fir, float *delta,
                       float *s_w, float *rms_out, int B, int T, int mode, float peak,
                       float max_rms, float eps, void *stream) {
  WM_ENTRY();
  WM_CHECK_ARG(B >= 0 && T >= 0, "postprocess: negative size");
  WM_CHECK_ARG(B == 0 || T == 0 || (delta_raw && (s || !s_w)), "postprocess: null pointer");
  return launch_postprocess(delta_raw, s, fir, delta, s_w, rms_out, B, T, mode, peak, max_rms, eps,
                            as_stream(stream));
}

int wm_detect_heads_fwd(const float *logits, const int *valid_len, float *probs, float *clip_prob,
                        float *msg_logits, float *vote_frac, int B, int T, int nout, void *stream) {
  WM_ENTRY();
  WM_CHECK_ARG(B >= 0 && T >= 0, "detect_heads: negative size");
  WM_CHECK_ARG(B == 0 || T == 0 || logits, "detect_heads: null pointer");
  return launch_detect_heads(logits, valid_len, probs, clip_prob, msg_logits, vote_frac, B, T, nout,
                             as_stream(stream));
}

size_t wm_generator_workspace_bytes(int B, int T) {
  if (B <= 0 || T <= 0) return 0;
  return 3 * act_bytes(B, T) + align256((size_t)B * 64 * sizeof(float));
}

int wm_generator_fwd(const float *blob, const float *embedding, int64_t emb_rows,
                     const int64_t *message, const float *s, float *delta_raw, void *workspace,
                     size_t workspace_bytes, int B, int T, void *stream) {
  WM_ENTRY();
  WM_CHECK_ARG(B >= 0 && T >= 0, "generator: negative size");
  if (B == 0 || T == 0) return 0;
  WM_CHECK_ARG(blob && s && delta_raw && workspace, "generator: null pointer");
  WM_CHECK_ARG(workspace_bytes >= wm_generator_workspace_bytes(B, T),
               "generator: workspace too small (%zu < %zu)", workspace_bytes,
               wm_generator_workspace_bytes(B, T));
  Ws ws{(char *)workspace, workspace_bytes};
  void *a0 = ws.take(act_bytes(B, T)), *a1 = ws.take(act_bytes(B, T)), *a2 = ws.take(act_bytes(B, T));
  float *emb = (float *)ws.take((size_t)B * 64 * 4);
  return generator_run(blob, embedding, emb_rows, message, s, delta_raw, a0, a1, a2, emb, B, T,
                       as_stream(stream));
}

size_t wm_detector_workspace_bytes(int B, int T) {
  if (B <= 0 || T <= 0) return 0;
  return 3 * act_bytes(B, T);
}

int wm_detector_fwd(const float *blob, const float *x, float *logits, void *workspace,
                    size_t workspace_bytes, int B, int T, int nout, void *stream) {
  WM_ENTRY();
  WM_CHECK_ARG(B >= 0 && T >= 0, "detector: negative size");
  WM_CHECK_ARG(nout >= 1 && nout <= WM_MAX_HEAD, "detector: nout must be in [1,%d]", WM_MAX_HEAD);
  if (B == 0 || T == 0) return 0;
  WM_CHECK_ARG(blob && x && logits && workspace, "detector: null pointer");
  WM_CHECK_ARG(workspace_bytes >= wm_detector_workspace_bytes(B, T), "detector: workspace too small");
  Ws ws{(char *)workspace, workspace_bytes};
  void *a0 = ws.take(act_bytes(B, T)), *a1 = ws.take(act_bytes(B, T)), *a2 = ws.take(act_bytes(B, T));
  float *out = nullptr;
  cudaStream_t st = as_stream(stream);
  WM_TRY(detector_trunk(blob, x, a0, a1, a2, &out, B, T, st));
  return launch_head(out, blob + WM_D_HEAD_W, blob + WM_D_HEAD_B, logits, B, T, nout, st);
}

int wm_detect_fwd(const float *blob, const float *x, const int *valid_len, float *probs,
                  float *clip_prob, float *msg_logits, float *vote_frac, void *workspace,
                  size_t workspace_bytes, int B, int T, int nout, void *stream) {
  WM_ENTRY();
  WM_CHECK_ARG(B >= 0 && T >= 0, "detect: negative size");
  WM_CHECK_ARG(nout >= 1 && nout <= WM_MAX_HEAD, "detect: nout must be in [1,%d]", WM_MAX_HEAD);
  if (B == 0 || T == 0) return 0;
  WM_CHECK_ARG(blob && x && workspace, "detect: null pointer");
  WM_CHECK_ARG(workspace_bytes >= wm_detector_workspace_bytes(B, T), "detect: workspace too small");
  Ws ws{(char *)workspace, workspace_bytes};
  void *a0 = ws.take(act_bytes(B, T)), *a1 = ws.take(act_bytes(B, T)), *a2 = ws.take(act_bytes(B, T));
  return detect_run(blob, x, valid_len, probs, clip_prob, msg_logits, vote_frac, a0, a1, a2, B, T, nout,
                    as_stream(stream));
}

size_t wm_embed_detect_workspace_bytes(int B, int T) {
  if (B <= 0 || T <= 0) return 0;
  return 3 * act_bytes(B, T) + align256((size_t)B * 64 * sizeof(float)) +
         align256((size_t)B * T * sizeof(float));
}

int wm_embed_detect_fwd(const float *g_blob, const float *embedding, int64_t emb_rows,
                        const float *d_blob, const float *fir, const int64_t *message,
                        const float *s, float *delta, float *s_w, float *delta_rms, float *probs,
                        float *clip_prob, float *msg_logits, float *vote_frac, void *workspace,
                        size_t workspace_bytes, int B, int T, int nout, int post_mode, void *stream) {
  WM_ENTRY();
  WM_CHECK_ARG(B >= 0 && T >= 0, "embed_detect: negative size");
  WM_CHECK_ARG(nout >= 1 && nout <= WM_MAX_HEAD, "embed_detect: nout must be in [1,%d]", WM_MAX_HEAD);
  if (B == 0 || T == 0) return 0;
  WM_CHECK_ARG(g_blob && d_blob && s && s_w && workspace, "embed_detect: null pointer");
  WM_CHECK_ARG(workspace_bytes >= wm_embed_detect_workspace_bytes(B, T),
               "embed_detect: workspace too small (%zu < %zu)", workspace_bytes,
               wm_embed_detect_workspace_bytes(B, T));
  Ws ws{(char *)workspace, workspace_bytes};
  void *a0 = ws.take(act_bytes(B, T)), *a1 = ws.take(act_bytes(B, T)), *a2 = ws.take(act_bytes(B, T));
  float *emb = (float *)ws.take((size_t)B * 64 * 4), *draw = (float *)ws.take((size_t)B * T * 4);
  cudaStream_t st = as_stream(stream);
  WM_TRY(generator_run(g_blob, embedding, emb_rows, message, s, draw, a0, a1, a2, emb, B, T, st));
  WM_TRY(launch_postprocess(draw, s, fir, delta, s_w, delta_rms, B, T, post_mode, 0.02f, 0.005f, 1e-8f, st));
  return detect_run(d_blob, s_w, nullptr, probs, clip_prob, msg_logits, vote_frac, a0, a1, a2, B, T, nout, st);
}

/* ---- training-loss forward (py/main16.py:74-81, 192-217, 255-266) ---- */
size_t wm_loss_workspace_bytes(int B, int T) {
  if (B <= 0 || T <= 0) return 0;
  return align256(((size_t)B * (1 + (size_t)T / 64) + 8192) * sizeof(float));
}

int wm_stft_frames(int T, int hop) { return (T <= 0 || hop <= 0) ? 0 : 1 + T / hop; }

int wm_stft_mag_fwd(const float *x, float *mag, int B, int T, int n_fft, int hop, void *stream) {
  WM_ENTRY();
  WM_CHECK_ARG(B >= 0 && T >= 0 && hop > 0, "stft_mag: bad size");
  if (B == 0 || T == 0) return 0;
  WM_CHECK_ARG(x && mag, "stft_mag: null pointer");
  return launch_stft_mag(x, mag, B, T, n_fft, hop, as_stream(stream));
}

#define WM_LOSS_ARGS(name)                                                                        \
  WM_ENTRY();                                                                                     \
  WM_CHECK_ARG(B > 0 && T > 0, name ": the mean over an empty batch is undefined");               \
  WM_CHECK_ARG(out && workspace, name ": null pointer");                                          \
  WM_CHECK_ARG(workspace_bytes >= wm_loss_workspace_bytes(B, T), name ": workspace too small")

int wm_hf_penalty_fwd(const float *delta, float *out, void *workspace, size_t workspace_bytes, int B, int T, int n_fft,
                      int first_bin, void *stream) {
  WM_LOSS_ARGS("hf_penalty");
  WM_CHECK_ARG(delta, "hf_penalty: null pointer");
  WM_CHECK_ARG(first_bin >= 0 && first_bin <= n_fft / 2 + 1, "hf_penalty: first_bin out of range");
  return launch_hf_penalty(delta, out, (float *)workspace, B, T, n_fft, first_bin, as_stream(stream));
}

int wm_loud_fwd(const float *clean, const float *wmk, float *out, void *workspace, size_t workspace_bytes, int B, int T,
                int n_fft, int hop, float thresh, void *stream) {
  WM_LOSS_ARGS("loud");
  WM_CHECK_ARG(clean && wmk && hop >= 64, "loud: null pointer or hop < 64");
  return launch_loudness(clean, wmk, out, (float *)workspace, B, T, n_fft, hop, thresh, as_stream(stream));
}

int wm_mel_log_l1_fwd(const float *clean, const float *wmk, const float *fb, const int *band, int n_mels, float *out,
                      void *workspace, size_t workspace_bytes, int B, int T, int n_fft, int hop, void *stream) {
  WM_LOSS_ARGS("mel_log_l1");
  WM_CHECK_ARG(clean && wmk && fb && band && n_mels > 0 && hop >= 64, "mel_log_l1: null pointer or bad size");
  return launch_mel_log_l1(clean, wmk, fb, band, n_mels, out, (float *)workspace, B, T, n_fft, hop, as_stream(stream));
}

int wm_bce_heads_fwd(const float *logits, const int64_t *message, float *loc_out, float *bce_out, void *workspace,
                     size_t workspace_bytes, int B_wm, int B_total, int T, int nout, void *stream) {
  WM_ENTRY();
  WM_CHECK_ARG(B_total > 0 && T > 0 && B_wm >= 0 && B_wm <= B_total, "bce_heads: bad batch sizes");
  WM_CHECK_ARG(nout >= 1 && nout <= WM_MAX_HEAD, "bce_heads: nout must be in [1,%d]", WM_MAX_HEAD);
  WM_CHECK_ARG(logits && loc_out && workspace && (message || nout == 1 || B_wm == 0), "bce_heads: null pointer");
  WM_CHECK_ARG(workspace_bytes >= wm_loss_workspace_bytes(B_total, T), "bce_heads: workspace too small");
  return launch_bce_heads(logits, message, loc_out, bce_out, (float *)workspace, B_wm, B_total, T, nout,
                          as_stream(stream));
}

int wm_abs_mean_fwd(const float *x, float *out, void *workspace, size_t workspace_bytes, int B, int T, void *stream) {
  WM_LOSS_ARGS("abs_mean");
  WM_CHECK_ARG(x, "abs_mean: null pointer");
  return launch_abs_mean(x, (long long)B * T, out, (float *)workspace, as_stream(stream));
}

/* ---- backward of the losses / post-processing ---- */
size_t wm_stft_bwd_workspace_bytes(int B, int T, int n_fft, int hop) {
  if (B <= 0 || T <= 0 || n_fft <= 0 || hop <= 0) return 0;
  return align256(stft_bwd_scratch_floats(B, T, n_fft, hop) * sizeof(float));
}

#define WM_LOSS_BWD_ARGS(name, hop_)                                                                       \
  WM_ENTRY();                                                                                              \
  WM_CHECK_ARG(B > 0 && T > 0, name ": the mean over an empty batch is undefined");                        \
  WM_CHECK_ARG(workspace && (hop_) >= 64, name ": null workspace or hop < 64");                            \
  WM_CHECK_ARG(workspace_bytes >= wm_stft_bwd_workspace_bytes(B, T, n_fft, (hop_)), name ": workspace too small")

int wm_hf_penalty_bwd(const float *delta, float *d_delta, void *workspace, size_t workspace_bytes, int B, int T,
                      int n_fft, int first_bin, float weight, int accumulate, void *stream) {
  WM_LOSS_BWD_ARGS("hf_penalty_bwd", n_fft / 4);
  WM_CHECK_ARG(delta && d_delta, "hf_penalty_bwd: null pointer");
  WM_CHECK_ARG(first_bin >= 0 && first_bin <= n_fft / 2 + 1, "hf_penalty_bwd: first_bin out of range");
  return launch_hf_penalty_bwd(delta, d_delta, (float *)workspace, B, T, n_fft, first_bin, weight, accumulate,
                               as_stream(stream));
}

int wm_loud_bwd(const float *clean, const float *wmk, float *d_wmk, void *workspace, size_t workspace_bytes, int B,
                int T, int n_fft, int hop, float thresh, float weight, int accumulate, void *stream) {
  WM_LOSS_BWD_ARGS("loud_bwd", hop);
  WM_CHECK_ARG(clean && wmk && d_wmk, "loud_bwd: null pointer");
  return launch_loudness_bwd(clean, wmk, d_wmk, (float *)workspace, B, T, n_fft, hop, thresh, weight, accumulate,
                             as_stream(stream));
}

int wm_mel_log_l1_bwd(const float *clean, const float *wmk, const float *fb, const int *band, int n_mels, float *d_wmk,
                      void *workspace, size_t workspace_bytes, int B, int T, int n_fft, int hop, float weight,
                      int accumulate, void *stream) {
  WM_LOSS_BWD_ARGS("mel_log_l1_bwd", hop);
  WM_CHECK_ARG(clean && wmk && fb && band && d_wmk && n_mels > 0 && n_mels <= 128, "mel_log_l1_bwd: bad arguments");
  return launch_mel_log_l1_bwd(clean, wmk, fb, band, n_mels, d_wmk, (float *)workspace, B, T, n_fft, hop, weight,
                               accumulate, as_stream(stream));
}

int wm_abs_mean_bwd(const float *x, float *dx, int B, int T, float weight, int accumulate, void *stream) {
  WM_ENTRY();
  WM_CHECK_ARG(B > 0 && T > 0 && x && dx, "abs_mean_bwd: bad arguments");
  return launch_abs_mean_bwd(x, dx, (long long)B * T, weight, accumulate, as_stream(stream));
}

int wm_postprocess_bwd(const float *g, const float *delta_fir, const float *fir, float *d_delta_raw, void *workspace,
                       size_t workspace_bytes, int B, int T, int mode, float peak, float max_rms, float eps,
                       void *stream) {
  WM_ENTRY();
  WM_CHECK_ARG(B >= 0 && T >= 0, "postprocess_bwd: negative size");
  if (B == 0 || T == 0) return 0;
  WM_CHECK_ARG(g && delta_fir && d_delta_raw && workspace && (fir || !(mode & 1)), "postprocess_bwd: null pointer");
  WM_CHECK_ARG(workspace_bytes >= (size_t)B * T * sizeof(float), "postprocess_bwd: workspace too small");
  return launch_postprocess_bwd(g, delta_fir, fir, d_delta_raw, (float *)workspace, B, T, mode, peak, max_rms, eps,
                                as_stream(stream));
}

/* ---- training (wm_train.cu) ---- */
size_t wm_detector_train_workspace_bytes(int B_total, int T, int nout) {
  if (B_total <= 0 || T <= 0 || nout < 1 || nout > WM_MAX_HEAD) return 0;
  return detector_train_workspace_bytes(B_total, T, nout);
}

int wm_detector_train_step(float *params, float *grads, float *adam_m, float *adam_v, float *run_stats, const float *x,
                           const int64_t *message, int B_wm, int B_total, int T, int nout, float lam_loc, float lam_dec,
                           float lr, float beta1, float beta2, float eps, int adam_step, float *losses_out,
                           float *d_input, void *workspace, size_t workspace_bytes, void *stream) {
  WM_ENTRY();
  WM_CHECK_ARG(B_total > 0 && T > 0 && B_wm >= 0 && B_wm <= B_total, "detector_train_step: bad batch sizes");
  WM_CHECK_ARG((long long)B_total * T > 1, "detector_train_step: BatchNorm needs more than one value per channel");
  WM_CHECK_ARG(nout >= 1 && nout <= WM_MAX_HEAD, "detector_train_step: nout must be in [1,%d]", WM_MAX_HEAD);
  WM_CHECK_ARG(params && grads && run_stats && x && workspace && (message || nout == 1 || B_wm == 0),
               "detector_train_step: null pointer");
  WM_CHECK_ARG(adam_step >= 0 && (adam_step == 0 || (adam_m && adam_v)), "detector_train_step: Adam state missing");
  WM_CHECK_ARG(workspace_bytes >= detector_train_workspace_bytes(B_total, T, nout),
               "detector_train_step: workspace too small");
  return detector_train_step(params, grads, adam_m, adam_v, run_stats, x, message, B_wm, B_total, T, nout, lam_loc,
                             lam_dec, lr, beta1, beta2, eps, adam_step, losses_out, d_input, workspace,
                             as_stream(stream));
}

size_t wm_train_step_workspace_bytes(int B, int T, int nout) {
  if (B <= 0 || T <= 0 || nout < 1 || nout > WM_MAX_HEAD) return 0;
  return train_step_workspace_bytes(B, T, nout);
}

int wm_train_forward_backward(const float *g_params, float *g_grads, float *g_stats, const float *d_params,
                              float *d_grads, float *d_stats, const float *s, const int64_t *message, const float *fir,
                              const float *mel_fb, const int *mel_band, int n_mels, const float *lam, int B, int T,
                              int nout, float *losses_out, float *s_w_out, void *workspace, size_t workspace_bytes,
                              void *stream) {
  WM_ENTRY();
  WM_CHECK_ARG(B > 0 && T > 0, "train_forward_backward: bad batch size");
  WM_CHECK_ARG(T > 1024, "train_forward_backward: the 2048-point STFT's reflect padding needs T > 1024 (T=%d)", T);
  WM_CHECK_ARG(nout >= 1 && nout <= WM_MAX_HEAD, "train_forward_backward: nout must be in [1,%d]", WM_MAX_HEAD);
  WM_CHECK_ARG(g_params && g_grads && g_stats && d_params && d_grads && d_stats && s && message && fir && mel_fb &&
                   mel_band && lam && workspace,
               "train_forward_backward: null pointer");
  WM_CHECK_ARG(n_mels > 0 && n_mels <= 128, "train_forward_backward: n_mels must be in [1,128]");
  WM_CHECK_ARG(workspace_bytes >= train_step_workspace_bytes(B, T, nout), "train_forward_backward: workspace too small");
  return train_forward_backward(g_params, g_grads, g_stats, d_params, d_grads, d_stats, s, message, fir, mel_fb, mel_band,
                                n_mels, lam, B, T, nout, losses_out, s_w_out, workspace, as_stream(stream));
}

int wm_bce_heads_bwd(const float *logits, const int64_t *message, float *dlogits, int B_wm, int B_total, int T, int nout,
                     float lam_loc, float lam_dec, void *stream) {
  WM_ENTRY();
  WM_CHECK_ARG(B_total > 0 && T > 0 && B_wm >= 0 && B_wm <= B_total, "bce_heads_bwd: bad batch sizes");
  WM_CHECK_ARG(nout >= 1 && nout <= WM_MAX_HEAD, "bce_heads_bwd: nout must be in [1,%d]", WM_MAX_HEAD);
  WM_CHECK_ARG(logits && dlogits && (message || nout == 1 || B_wm == 0), "bce_heads_bwd: null pointer");
  return launch_bce_heads_bwd(logits, message, B_wm, B_total, T, nout, lam_loc, lam_dec, dlogits, as_stream(stream));
}

size_t wm_head_bwd_workspace_bytes(long long rows, int nout) {
  return rows > 0 && nout > 0 ? head_bwd_scratch_floats(rows, nout) * sizeof(float) : 0;
}

int wm_head_bwd(const float *dlogits, const float *y, const float *w, float *dy, float *dw, float *db, long long rows,
                int nout, void *workspace, size_t workspace_bytes, void *stream) {
  WM_ENTRY();
  WM_CHECK_ARG(rows > 0 && nout >= 1 && nout <= WM_MAX_HEAD, "head_bwd: bad size");
  WM_CHECK_ARG(dlogits && y && w && dy && dw && db && workspace, "head_bwd: null pointer");
  WM_CHECK_ARG(workspace_bytes >= wm_head_bwd_workspace_bytes(rows, nout), "head_bwd: workspace too small");
  return launch_head_bwd(dlogits, y, w, dy, dw, db, rows, nout, (float *)workspace, as_stream(stream));
}

size_t wm_conv_in_k7_bwd_workspace_bytes(int B, int T) {
  return B > 0 && T > 0 ? conv_in_grads_scratch_floats(B, T) * sizeof(float) : 0;
}

int wm_conv_in_k7_bwd(const float *s, const float *dx, const float *w, float *dw, float *db, float *ds, int B, int T,
                      void *workspace, size_t workspace_bytes, void *stream) {
  WM_ENTRY();
  WM_CHECK_ARG(B > 0 && T > 0, "conv_in_k7_bwd: bad size");
  WM_CHECK_ARG(s && dx && dw && db && workspace && (!ds || w), "conv_in_k7_bwd: null pointer");
  WM_CHECK_ARG(workspace_bytes >= wm_conv_in_k7_bwd_workspace_bytes(B, T), "conv_in_k7_bwd: workspace too small");
  return launch_conv_in_grads(s, dx, w, dw, db, ds, B, T, (float *)workspace, as_stream(stream));
}

size_t wm_bn_train_workspace_bytes(long long rows) { return rows > 0 ? train_scratch_doubles(rows) * sizeof(double) : 0; }

int wm_bn_train_fwd(const float *z, const float *gamma, const float *beta, const float *residual, float *out,
                    float *mean, float *rstd, float *run_mean, float *run_var, long long rows, int relu,
                    void *workspace, size_t workspace_bytes, void *stream) {
  WM_ENTRY();
  WM_CHECK_ARG(rows > 1, "bn_train_fwd: needs more than one row");
  WM_CHECK_ARG(z && gamma && beta && out && mean && rstd && workspace && (!run_mean == !run_var),
               "bn_train_fwd: null pointer");
  WM_CHECK_ARG(workspace_bytes >= wm_bn_train_workspace_bytes(rows), "bn_train_fwd: workspace too small");
  return launch_bn_train_fwd(z, gamma, beta, residual, out, mean, rstd, run_mean, run_var, rows, relu,
                             (double *)workspace, as_stream(stream));
}

int wm_bn_train_bwd(const float *dout, const float *act, const float *z, const float *mean, const float *rstd,
                    const float *gamma, float *dz, float *dres, float *dgamma, float *dbeta, long long rows,
                    void *workspace, size_t workspace_bytes, void *stream) {
  WM_ENTRY();
  WM_CHECK_ARG(rows > 1, "bn_train_bwd: needs more than one row");
  WM_CHECK_ARG(dout && z && mean && rstd && gamma && dz && dgamma && dbeta && workspace, "bn_train_bwd: null pointer");
  WM_CHECK_ARG(workspace_bytes >= wm_bn_train_workspace_bytes(rows), "bn_train_bwd: workspace too small");
  return launch_bn_train_bwd(dout, act, z, mean, rstd, gamma, dz, dres, dgamma, dbeta, rows, (double *)workspace,
                             as_stream(stream));
}

size_t wm_conv64_bwd_workspace_bytes(int B, int T, int K) {
  if (B <= 0 || T <= 0 || K <= 0) return 0;
  return train_conv64_scratch_floats(B, T, K) * sizeof(float);
}

int wm_conv64_bwd(const float *x, const float *dy, const float *w, float *dw, float *db, float *dx, int B, int T, int K,
                  void *workspace, size_t workspace_bytes, void *stream) {
  WM_ENTRY();
  WM_CHECK_ARG(B > 0 && T > 0, "conv64_bwd: bad size");
  WM_CHECK_ARG(K == 1 || K == 3 || K == 7, "conv64_bwd: K must be 1, 3 or 7 (got %d)", K);
  WM_CHECK_ARG(x && dy && dw && db && workspace && (!dx || w), "conv64_bwd: null pointer");
  WM_CHECK_ARG(workspace_bytes >= wm_conv64_bwd_workspace_bytes(B, T, K), "conv64_bwd: workspace too small");
  return train_conv64_bwd(x, dy, w, dw, db, dx, B, T, K, (float *)workspace, as_stream(stream));
}

int wm_conv64_train_fwd(const float *x, const float *w, const float *bias, const float *residual, float *y, int B, int T,
                        int K, void *workspace, size_t workspace_bytes, void *stream) {
  WM_ENTRY();
  WM_CHECK_ARG(B >= 0 && T >= 0, "conv64_train_fwd: negative size");
  if (B == 0 || T == 0) return 0;
  WM_CHECK_ARG(K == 1 || K == 3 || K == 7, "conv64_train_fwd: K must be 1, 3 or 7 (got %d)", K);
  WM_CHECK_ARG(x && w && bias && y && workspace && x != y, "conv64_train_fwd: null pointer or in-place");
  WM_CHECK_ARG(workspace_bytes >= wm_conv64_bwd_workspace_bytes(B, T, K), "conv64_train_fwd: workspace too small");
  return train_conv64_fwd(x, w, bias, residual, y, B, T, K, (float *)workspace, as_stream(stream));
}

int wm_lstm_train_fwd(const float *x, const float *wT_ih, const float *wT_hh, const float *b_ih, const float *b_hh,
                      float *h, float *gates, float *cell, int B, int T, void *stream) {
  WM_ENTRY();
  WM_CHECK_ARG(B >= 0 && T >= 0, "lstm_train_fwd: bad size");
  WM_CHECK_ARG(x && wT_ih && wT_hh && b_ih && b_hh && h && gates && cell, "lstm_train_fwd: null pointer");
  return launch_lstm_train_fwd(x, wT_ih, wT_hh, b_ih, b_hh, h, gates, cell, B, T, as_stream(stream));
}

size_t wm_lstm_train_bwd_workspace_bytes(int B, int T) {
  return B > 0 && T > 0 ? lstm_train_bwd_scratch_floats(B, T) * sizeof(float) : 0;
}

int wm_lstm_train_bwd(const float *dy, const float *x, const float *h, const float *wT_ih, const float *wT_hh,
                      const float *gates, const float *cell, float *dx, float *dwT_ih, float *dwT_hh, float *db, int B,
                      int T, void *workspace, size_t workspace_bytes, void *stream) {
  WM_ENTRY();
  WM_CHECK_ARG(B > 0 && T > 0, "lstm_train_bwd: bad size");
  WM_CHECK_ARG(dy && x && h && wT_ih && wT_hh && gates && cell && dx && dwT_ih && dwT_hh && db && workspace,
               "lstm_train_bwd: null pointer");
  WM_CHECK_ARG(workspace_bytes >= wm_lstm_train_bwd_workspace_bytes(B, T), "lstm_train_bwd: workspace too small");
  return launch_lstm_train_bwd(dy, x, h, wT_ih, wT_hh, gates, cell, dx, dwT_ih, dwT_hh, db, B, T, (float *)workspace,
                               as_stream(stream));
}

int wm_adam_step(float *p, const float *g, float *m, float *v, long long n, float lr, float beta1, float beta2,
                 float eps, int step, void *stream) {
  WM_ENTRY();
  WM_CHECK_ARG(p && g && m && v && n >= 0 && step >= 1, "adam_step: bad arguments");
  if (n == 0) return 0;
  return launch_adam(p, g, m, v, n, lr, beta1, beta2, eps, step, as_stream(stream));
}

/* ---- generic operators of the main14b_2 stack (py/main14b_2.py:83-224) ---- */
int wm_conv1d_out_len(int Tin, int K, int stride, int pad) { return (Tin + 2 * pad - K) / stride + 1; }
int wm_convtranspose1d_out_len(int Tin, int K, int stride, int pad) { return (Tin - 1) * stride - 2 * pad + K; }

int wm_conv1d_fwd(const float *x, const float *w, const float *bias, const float *chan_add, const float *residual,
                  float *y, int B, int Cin, int Tin, int Cout, int K, int stride, int pad, int act, void *stream) {
  WM_ENTRY();
  WM_CHECK_ARG(B >= 0 && Cin > 0 && Cout > 0 && Tin >= 0, "conv1d: bad size");
  WM_CHECK_ARG(act == 0 || act == 1, "conv1d: act must be 0 (none) or 1 (ELU)");
  if (B == 0 || Tin == 0) return 0;
  WM_CHECK_ARG(x && w && bias && y, "conv1d: null pointer");
  WM_CHECK_ARG(x != y && residual != y, "conv1d: in-place operation is not supported");
  return launch_conv1d_generic(x, w, bias, chan_add, residual, y, B, Cin, Tin, Cout, K, stride, pad, act,
                               as_stream(stream));
}

int wm_convtranspose1d_fwd(const float *x, const float *w, const float *bias, float *y, int B, int Cin, int Tin,
                           int Cout, int K, int stride, int pad, void *stream) {
  WM_ENTRY();
  WM_CHECK_ARG(B >= 0 && Cin > 0 && Cout > 0 && Tin >= 0, "convtranspose1d: bad size");
  if (B == 0 || Tin == 0) return 0;
  WM_CHECK_ARG(x && w && bias && y && x != y, "convtranspose1d: null pointer or in-place");
  return launch_convtranspose1d_generic(x, w, bias, y, B, Cin, Tin, Cout, K, stride, pad, as_stream(stream));
}

size_t wm_convtranspose1d_phase_weight_floats(int Cin, int Cout, int stride) {
  return (Cin <= 0 || Cout <= 0 || stride <= 0) ? 0 : (size_t)Cout * stride * Cin * 3 + (size_t)Cout * stride;
}

int wm_convtranspose1d_pack(const float *w, const float *bias, float *packed, int Cin, int Cout, int K, int stride,
                            int pad, void *stream) {
  WM_ENTRY();
  WM_CHECK_ARG(w && bias && packed, "convtranspose1d_pack: null pointer");
  WM_CHECK_ARG(K == 2 * stride && pad >= 0 && pad < stride && stride >= 1 && stride <= 8,
               "convtranspose1d_pack: the phase form needs kernel_size == 2 * stride and padding < stride");
  return launch_convt_phase_weights(w, bias, packed, packed + (size_t)Cout * stride * Cin * 3, Cin, Cout, stride, pad,
                                    as_stream(stream));
}

int wm_convtranspose1d_phase_fwd(const float *x, const float *packed, float *y, int B, int Cin, int Tin, int Cout, int K,
                                 int stride, int pad, void *stream) {
  WM_ENTRY();
  WM_CHECK_ARG(B >= 0 && Cin > 0 && Cout > 0 && Tin >= 0, "convtranspose1d_phase: bad size");
  WM_CHECK_ARG(K == 2 * stride && pad >= 0 && pad < stride, "convtranspose1d_phase: needs kernel_size == 2 * stride");
  if (B == 0 || Tin == 0) return 0;
  WM_CHECK_ARG(x && packed && y && x != y, "convtranspose1d_phase: null pointer or in-place");
  const int Tout = (Tin - 1) * stride - 2 * pad + K;
  const int extra = (Tout + stride - 1) / stride - Tin;       // phase-0 samples beyond the input length
  return launch_conv1d_generic(x, packed, packed + (size_t)Cout * stride * Cin * 3, nullptr, nullptr, y, B, Cin, Tin,
                               Cout * stride, 3, 1, 1, 0, as_stream(stream), stride, Tout, extra > 0 ? extra : 0);
}

int wm_lstm_small_fwd(const float *x, const float *w_ih, const float *w_hh, const float *bias, float *y, int B, int H,
                      int T, int layers, void *stream) {
  WM_ENTRY();
  WM_CHECK_ARG(B >= 0 && T >= 0, "lstm_small: negative size");
  if (B == 0 || T == 0) return 0;
  WM_CHECK_ARG(x && w_ih && w_hh && bias && y && x != y, "lstm_small: null pointer or in-place");
  return launch_lstm_small(x, w_ih, w_hh, bias, y, B, H, T, layers, as_stream(stream));
}

/* ---- audio formats either side of the path ---- */
int wm_resample_fwd(const float *x, const float *kern, float *y, int B, int Tin, int Tout, int down, int up, int K,
                    int width, void *stream) {
  WM_ENTRY();
  WM_CHECK_ARG(B >= 0 && Tin >= 0 && Tout >= 0 && down > 0 && up > 0 && K > 0 && width >= 0, "resample: bad size");
  if (B == 0 || Tout == 0) return 0;
  WM_CHECK_ARG(x && kern && y && x != y, "resample: null pointer or in-place");
  return launch_resample(x, kern, y, B, Tin, Tout, down, up, K, width, as_stream(stream));
}

int wm_pcm16_quantize_fwd(const float *x, int16_t *q, size_t n, void *stream) {
  WM_ENTRY();
  WM_CHECK_ARG(n == 0 || (x && q), "pcm16_quantize: null pointer");
  return launch_pcm16(x, reinterpret_cast<short *>(q), nullptr, (long long)n, 1, 0.0f, as_stream(stream));
}

int wm_pcm16_dequantize_fwd(const int16_t *q, float *x, size_t n, float scale, void *stream) {
  WM_ENTRY();
  WM_CHECK_ARG(n == 0 || (x && q), "pcm16_dequantize: null pointer");
  return launch_pcm16(nullptr, reinterpret_cast<short *>(const_cast<int16_t *>(q)), x, (long long)n, 0, scale,
                      as_stream(stream));
}

int wm_file_metrics_fwd(const float *s, const float *s_w, const int *valid_len, float *out, int B, int T, void *stream) {
  WM_ENTRY();
  WM_CHECK_ARG(B >= 0 && T >= 0, "file_metrics: negative size");
  if (B == 0) return 0;
  WM_CHECK_ARG(s && s_w && out, "file_metrics: null pointer");
  return launch_file_metrics(s, s_w, valid_len, out, B, T, as_stream(stream));
}

size_t wm_biquad_workspace_bytes(int rows, long long N) { return rows > 0 && N > 0 ? biquad_scratch_bytes(rows, N) : 0; }

int wm_biquad_fwd(const float *x, float *y, int16_t *pcm16, int rows, long long N, const double *b3, const double *a3,
                  int clamp, void *workspace, size_t workspace_bytes, void *stream) {
  WM_ENTRY();
  WM_CHECK_ARG(rows >= 0 && N >= 0, "biquad: negative size");
  if (rows == 0 || N == 0) return 0;
  WM_CHECK_ARG(x && y && b3 && a3 && workspace, "biquad: null pointer");
  WM_CHECK_ARG(x != y, "biquad: in-place operation is not supported");
  WM_CHECK_ARG(a3[0] != 0.0, "biquad: a[0] must not be zero");
  WM_CHECK_ARG(workspace_bytes >= biquad_scratch_bytes(rows, N), "biquad: workspace too small");
  return launch_biquad(x, y, reinterpret_cast<short *>(pcm16), rows, N, b3, a3, clamp, workspace, as_stream(stream));
}

int wm_confusion_counts_fwd(const float *clean, long long n_clean, const float *wm, long long n_wm, float thresh,
                            unsigned long long *out4, void *stream) {
  WM_ENTRY();
  WM_CHECK_ARG(n_clean >= 0 && n_wm >= 0 && out4 && (clean || n_clean == 0) && (wm || n_wm == 0), "confusion_counts: bad arguments");
  return launch_confusion(clean, n_clean, wm, n_wm, thresh, out4, as_stream(stream));
}

int wm_roc_points_fwd(const float *clean, long long n_clean, const float *wm, long long n_wm, const float *thresholds,
                      int n_thresholds, int *fp, int *tp, void *stream) {
  WM_ENTRY();
  WM_CHECK_ARG(n_clean >= 0 && n_wm >= 0 && n_thresholds >= 0, "roc_points: negative size");
  WM_CHECK_ARG(n_clean < (1LL << 31) && n_wm < (1LL << 31), "roc_points: counts are 32-bit");
  if (n_thresholds == 0) return 0;
  WM_CHECK_ARG(thresholds && fp && tp && (clean || n_clean == 0) && (wm || n_wm == 0), "roc_points: null pointer");
  return launch_roc_points(clean, n_clean, wm, n_wm, thresholds, n_thresholds, fp, tp, as_stream(stream));
}

int wm_auc_pairs_fwd(const float *clean, long long n_clean, const float *wm, long long n_wm, unsigned long long *out,
                     void *stream) {
  WM_ENTRY();
  WM_CHECK_ARG(n_clean >= 0 && n_wm >= 0 && out && (clean || n_clean == 0) && (wm || n_wm == 0), "auc_pairs: bad arguments");
  return launch_auc_pairs(clean, n_clean, wm, n_wm, out, as_stream(stream));
}

size_t wm_embed_detect_host_workspace_bytes(int chunk, int T, int nout) {
  if (chunk <= 0 || T <= 0 || nout < 1) return 0;
  size_t wave = align256((size_t)chunk * T * sizeof(float));
  // two staging sets (chunk parity): s, s_w, probs, message, clip_prob, msg_logits
  return wm_embed_detect_workspace_bytes(chunk, T) +
         2 * (3 * wave + align256((size_t)chunk * sizeof(int64_t)) + align256((size_t)chunk * sizeof(float)) +
              align256((size_t)chunk * (nout - 1 > 0 ? nout - 1 : 1) * sizeof(float)));
}

namespace {

// Copy streams and events of the host-fed pipeline, one set per calling host thread (created on first use,
// kept for the life of the thread) so that concurrent callers on distinct streams never share an event.
constexpr int HP_NS = 4;   // sub-batches per device pass
struct HostPipe {
  int dev = -1;
  cudaStream_t in = nullptr, out = nullptr;
  cudaEvent_t start = nullptr, in_done[2][HP_NS], sw_done[2][HP_NS], pr_done[2][HP_NS], s_free[2], out_done[2];
  int init() {
    int d = 0;
    WM_CHECK_CUDA(cudaGetDevice(&d));
    if (dev == d) return 0;
    WM_CHECK_CUDA(cudaStreamCreateWithFlags(&in, cudaStreamNonBlocking));
    WM_CHECK_CUDA(cudaStreamCreateWithFlags(&out, cudaStreamNonBlocking));
    auto mk = [](cudaEvent_t *e) { return cudaEventCreateWithFlags(e, cudaEventDisableTiming); };
    WM_CHECK_CUDA(mk(&start));
    for (int p = 0; p < 2; ++p) {
      WM_CHECK_CUDA(mk(&s_free[p]));
      WM_CHECK_CUDA(mk(&out_done[p]));
      for (int k = 0; k < HP_NS; ++k) {
        WM_CHECK_CUDA(mk(&in_done[p][k]));
        WM_CHECK_CUDA(mk(&sw_done[p][k]));
        WM_CHECK_CUDA(mk(&pr_done[p][k]));
      }
    }
    dev = d;
    return 0;
  }
};
thread_local HostPipe g_pipe;

}  // namespace

// Host-fed embed+detect.  tcgen05 mode: every device pass of <= chunk clips is cut into HP_NS sub-batches;
// the convolutional phases (encoder; decoder + post-processing + detector) run per sub-batch so that the
// H2D copy of sub-batch k+1 and the D2H copies of sub-batch k-1 overlap them on two copy streams, while the
// latency-bound LSTM runs once over the whole pass.  Staging buffers alternate between passes, so pass c+1's
// input copy and pass c-1's output copy also overlap pass c.  On return, `stream` has been made to wait for
// every copy: stream-order semantics are those of a plain sequence of copies and kernels on `stream`.
// H2D of `n` floats of clip data starting `off` floats into the caller's buffer, of which only the first `valid` floats
// exist: the rest is zero on the device (a ragged last segment read straight out of the caller's pinned recording).
static int copy_in_ragged(float *dst, const float *host_s, long long off, long long n, long long valid, cudaStream_t st) {
  long long have = valid - off;
  have = have < 0 ? 0 : (have > n ? n : have);
  if (have > 0) WM_CHECK_CUDA(cudaMemcpyAsync(dst, host_s + off, (size_t)have * 4, cudaMemcpyHostToDevice, st));
  if (have < n) WM_CHECK_CUDA(cudaMemsetAsync(dst + have, 0, (size_t)(n - have) * 4, st));
  return 0;
}

int wm_embed_detect_host(const float *g_blob, const float *embedding, int64_t emb_rows,
                         const float *d_blob, const float *fir, const int64_t *host_message,
                         const float *host_s, float *host_s_w, float *host_probs,
                         float *host_clip_prob, float *host_msg_logits, void *workspace,
                         size_t workspace_bytes, int B, int T, int nout, int chunk, int post_mode,
                         void *stream) {
  return wm_embed_detect_host_ragged(g_blob, embedding, emb_rows, d_blob, fir, host_message, host_s, (long long)B * T,
                                     host_s_w, host_probs, host_clip_prob, host_msg_logits, workspace, workspace_bytes, B,
                                     T, nout, chunk, post_mode, stream);
}

int wm_embed_detect_host_ragged(const float *g_blob, const float *embedding, int64_t emb_rows,
                                const float *d_blob, const float *fir, const int64_t *host_message,
                                const float *host_s, long long host_s_floats, float *host_s_w, float *host_probs,
                                float *host_clip_prob, float *host_msg_logits, void *workspace,
                                size_t workspace_bytes, int B, int T, int nout, int chunk, int post_mode,
                                void *stream) {
  WM_ENTRY();
  WM_CHECK_ARG(host_s_floats >= 0 && host_s_floats <= (long long)B * T,
               "embed_detect_host: host_s_floats %lld outside [0, B * T]", host_s_floats);
  WM_CHECK_ARG(B >= 0 && T >= 0 && chunk > 0, "embed_detect_host: bad size");
  WM_CHECK_ARG(nout >= 1 && nout <= WM_MAX_HEAD, "embed_detect_host: nout must be in [1,%d]", WM_MAX_HEAD);
  if (B == 0 || T == 0) return 0;
  WM_CHECK_ARG(g_blob && d_blob && host_s && host_s_w && workspace, "embed_detect_host: null pointer");
  WM_CHECK_ARG(workspace_bytes >= wm_embed_detect_host_workspace_bytes(chunk, T, nout),
               "embed_detect_host: workspace too small");
  const int nbits = nout - 1;
  Ws ws{(char *)workspace, workspace_bytes};
  const size_t wave = (size_t)chunk * T * sizeof(float);
  float *d_s[2], *d_sw[2], *d_pr[2], *d_cp[2], *d_ml[2];
  int64_t *d_msg[2];
  for (int p = 0; p < 2; ++p) {
    d_s[p] = (float *)ws.take(wave); d_sw[p] = (float *)ws.take(wave); d_pr[p] = (float *)ws.take(wave);
    d_msg[p] = (int64_t *)ws.take((size_t)chunk * sizeof(int64_t));
    d_cp[p] = (float *)ws.take((size_t)chunk * sizeof(float));
    d_ml[p] = (float *)ws.take((size_t)chunk * (nbits > 0 ? nbits : 1) * sizeof(float));
  }
  void *inner = ws.p;
  size_t inner_bytes = ws.left;
  cudaStream_t st = as_stream(stream);

  if (g_math_mode.load() != WM_MATH_BF16X2) {   // fp32 cross-check mode: plain sequence on `stream`
    for (int b0 = 0; b0 < B; b0 += chunk) {
      int nb = B - b0 < chunk ? B - b0 : chunk;
      WM_TRY(copy_in_ragged(d_s[0], host_s, (long long)b0 * T, (long long)nb * T, host_s_floats, st));
      if (host_message)
        WM_CHECK_CUDA(cudaMemcpyAsync(d_msg[0], host_message + b0, (size_t)nb * 8, cudaMemcpyHostToDevice, st));
      WM_TRY(wm_embed_detect_fwd(g_blob, embedding, emb_rows, d_blob, fir, host_message ? d_msg[0] : nullptr, d_s[0],
                                 nullptr, d_sw[0], nullptr, host_probs ? d_pr[0] : nullptr, d_cp[0],
                                 nbits > 0 ? d_ml[0] : nullptr, nullptr, inner, inner_bytes, nb, T, nout,
                                 post_mode, stream));
      WM_CHECK_CUDA(cudaMemcpyAsync(host_s_w + (size_t)b0 * T, d_sw[0], (size_t)nb * T * 4, cudaMemcpyDeviceToHost, st));
      if (host_probs)
        WM_CHECK_CUDA(cudaMemcpyAsync(host_probs + (size_t)b0 * T, d_pr[0], (size_t)nb * T * 4, cudaMemcpyDeviceToHost, st));
      if (host_clip_prob)
        WM_CHECK_CUDA(cudaMemcpyAsync(host_clip_prob + b0, d_cp[0], (size_t)nb * 4, cudaMemcpyDeviceToHost, st));
      if (host_msg_logits && nbits > 0)
        WM_CHECK_CUDA(cudaMemcpyAsync(host_msg_logits + (size_t)b0 * nbits, d_ml[0], (size_t)nb * nbits * 4,
                                      cudaMemcpyDeviceToHost, st));
    }
    return 0;
  }

  HostPipe &hp = g_pipe;
  WM_TRY(hp.init());
  Ws iw{(char *)inner, inner_bytes};
  char *a0 = (char *)iw.take(act_bytes(chunk, T)), *a1 = (char *)iw.take(act_bytes(chunk, T)),
       *a2 = (char *)iw.take(act_bytes(chunk, T));
  float *emb = (float *)iw.take((size_t)chunk * 64 * 4), *draw = (float *)iw.take((size_t)chunk * T * 4);
  WM_CHECK_ARG(a0 && a1 && a2 && emb && draw, "embed_detect_host: workspace too small");
  const size_t clip_bytes = (size_t)16 * ((size_t)T + 2 * WM_PLANAR_PAD) * 16;   // one clip of a planar buffer
  const bool use_msg = host_message != nullptr && embedding != nullptr;

  WM_CHECK_CUDA(cudaEventRecord(hp.start, st));
  WM_CHECK_CUDA(cudaStreamWaitEvent(hp.in, hp.start, 0));
  WM_CHECK_CUDA(cudaStreamWaitEvent(hp.out, hp.start, 0));
  int pass = 0;
  for (int b0 = 0; b0 < B; b0 += chunk, ++pass) {
    const int nb = B - b0 < chunk ? B - b0 : chunk;
    const int p = pass & 1;
    const int ns = nb >= 128 * HP_NS ? HP_NS : 1;
    const int sb = (nb + ns - 1) / ns;
    // ---- input copies (copy-in stream) ----
    if (pass >= 2) WM_CHECK_CUDA(cudaStreamWaitEvent(hp.in, hp.s_free[p], 0));
    if (use_msg)
      WM_CHECK_CUDA(cudaMemcpyAsync(d_msg[p], host_message + b0, (size_t)nb * 8, cudaMemcpyHostToDevice, hp.in));
    for (int k = 0; k < ns; ++k) {
      const int k0 = k * sb, kn = nb - k0 < sb ? nb - k0 : sb;
      if (kn <= 0) { WM_CHECK_CUDA(cudaEventRecord(hp.in_done[p][k], hp.in)); continue; }
      WM_TRY(copy_in_ragged(d_s[p] + (size_t)k0 * T, host_s, ((long long)b0 + k0) * T, (long long)kn * T, host_s_floats,
                            hp.in));
      WM_CHECK_CUDA(cudaEventRecord(hp.in_done[p][k], hp.in));
    }
    // ---- encoder per sub-batch, LSTM over the pass ----
    for (int k = 0; k < ns; ++k) {
      const int k0 = k * sb, kn = nb - k0 < sb ? nb - k0 : sb;
      WM_CHECK_CUDA(cudaStreamWaitEvent(st, hp.in_done[p][k], 0));
      if (kn <= 0) continue;
      WM_TRY(generator_encoder_tc(g_blob, d_s[p] + (size_t)k0 * T, a0 + k0 * clip_bytes, a1 + k0 * clip_bytes,
                                  a2 + k0 * clip_bytes, kn, T, st));
    }
    if (use_msg) WM_TRY(launch_gather_rows(embedding, emb_rows, d_msg[p], emb, nb, st));
    WM_TRY(generator_lstm_tc(g_blob, use_msg ? emb : nullptr, a1, a0, nb, T, st));
    // ---- decoder, post-processing, detector per sub-batch; outputs leave on the copy-out stream ----
    if (pass >= 2) WM_CHECK_CUDA(cudaStreamWaitEvent(st, hp.out_done[p], 0));   // staging set p is being reused
    for (int k = 0; k < ns; ++k) {
      const int k0 = k * sb, kn = nb - k0 < sb ? nb - k0 : sb;
      if (kn <= 0) continue;
      float *sk = d_s[p] + (size_t)k0 * T, *swk = d_sw[p] + (size_t)k0 * T, *prk = d_pr[p] + (size_t)k0 * T;
      WM_TRY(generator_decoder_tc(g_blob, a0 + k0 * clip_bytes, a1 + k0 * clip_bytes, draw + (size_t)k0 * T, kn, T, st));
      WM_TRY(launch_postprocess(draw + (size_t)k0 * T, sk, fir, nullptr, swk, nullptr, kn, T, post_mode, 0.02f, 0.005f,
                                1e-8f, st));
      WM_CHECK_CUDA(cudaEventRecord(hp.sw_done[p][k], st));
      WM_CHECK_CUDA(cudaStreamWaitEvent(hp.out, hp.sw_done[p][k], 0));
      WM_CHECK_CUDA(cudaMemcpyAsync(host_s_w + ((size_t)b0 + k0) * T, swk, (size_t)kn * T * 4, cudaMemcpyDeviceToHost,
                                    hp.out));
      WM_TRY(detect_run(d_blob, swk, nullptr, host_probs ? prk : nullptr, d_cp[p] + k0,
                        nbits > 0 ? d_ml[p] + (size_t)k0 * nbits : nullptr, nullptr, a0 + k0 * clip_bytes,
                        a1 + k0 * clip_bytes, a2 + k0 * clip_bytes, kn, T, nout, st));
      WM_CHECK_CUDA(cudaEventRecord(hp.pr_done[p][k], st));
      WM_CHECK_CUDA(cudaStreamWaitEvent(hp.out, hp.pr_done[p][k], 0));
      if (host_probs)
        WM_CHECK_CUDA(cudaMemcpyAsync(host_probs + ((size_t)b0 + k0) * T, prk, (size_t)kn * T * 4,
                                      cudaMemcpyDeviceToHost, hp.out));
    }
    WM_CHECK_CUDA(cudaEventRecord(hp.s_free[p], st));
    if (host_clip_prob)
      WM_CHECK_CUDA(cudaMemcpyAsync(host_clip_prob + b0, d_cp[p], (size_t)nb * 4, cudaMemcpyDeviceToHost, hp.out));
    if (host_msg_logits && nbits > 0)
      WM_CHECK_CUDA(cudaMemcpyAsync(host_msg_logits + (size_t)b0 * nbits, d_ml[p], (size_t)nb * nbits * 4,
                                    cudaMemcpyDeviceToHost, hp.out));
    WM_CHECK_CUDA(cudaEventRecord(hp.out_done[p], hp.out));
  }
  // `stream` completes only after every output has reached the host
  WM_CHECK_CUDA(cudaStreamWaitEvent(st, hp.out_done[(pass - 1) & 1], 0));
  if (pass >= 2) WM_CHECK_CUDA(cudaStreamWaitEvent(st, hp.out_done[pass & 1], 0));
  return 0;
}

}  // extern "C"
