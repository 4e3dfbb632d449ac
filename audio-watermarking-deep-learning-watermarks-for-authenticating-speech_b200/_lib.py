"""ctypes binding of libwmb200.so — the C ABI declared in include/wmb200.h.

This is the stub a maintainer of the reference would add: every
function takes raw device pointers, sizes and a CUDA stream handle.  There is no
fallback: if the library is missing or the device is not sm_100 the import of the
compute path raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("WMB200_LIB") or os.path.join(HERE, "libwmb200.so")   # WMB200_LIB: developer A/B builds
ABI_VERSION = 21

# blob offsets (floats) — mirror of the enums in include/wmb200.h
RB_W1 = 0
RB_B1 = RB_W1 + 3 * 64 * 64
RB_W2 = RB_B1 + 64
RB_B2 = RB_W2 + 3 * 64 * 64
RB_SIZE = RB_B2 + 64
FIN_W9 = 0
FIN_B9 = FIN_W9 + 9 * 64
FIN_WK = FIN_B9 + 64
FIN_BK = FIN_WK + 3 * 7 * 64
FIN_SIZE = FIN_BK + 3 * 64
G_IN_W = 0
G_IN_B = G_IN_W + 7 * 64
G_RB0 = G_IN_B + 64
G_RB1 = G_RB0 + RB_SIZE
G_LSTM_WIH = G_RB1 + RB_SIZE
G_LSTM_WHH = G_LSTM_WIH + 256 * 64
G_LSTM_B = G_LSTM_WHH + 256 * 64
G_CT_W = G_LSTM_B + 256
G_CT_B = G_CT_W + 7 * 64 * 64
G_RB2 = G_CT_B + 64
G_HEAD_W = G_RB2 + RB_SIZE
G_HEAD_B = G_HEAD_W + 64
G_FIN = G_HEAD_B + 4
G_SIZE = G_FIN + FIN_SIZE
TC_IMG3 = 3 * 8 * 128 * 8 // 2
TC_IMG7 = 7 * 8 * 128 * 8 // 2
G_TC = (G_SIZE + 63) // 64 * 64
G_TC_CT = G_TC + 4 * TC_IMG3
G_TC_RB2 = G_TC_CT + TC_IMG7
G_TC_LSTM_W = G_TC_RB2 + 2 * TC_IMG3
G_TC_LSTM_B = G_TC_LSTM_W + 4 * 256 * 64 // 2
G_BLOB = G_TC_LSTM_B + 256
D_IN_W = 0
D_IN_B = D_IN_W + 7 * 64
D_RB0 = D_IN_B + 64
D_RB1 = D_RB0 + RB_SIZE
D_HEAD_W = D_RB1 + RB_SIZE
D_HEAD_B = D_HEAD_W + 32 * 64
D_FIN = D_HEAD_B + 32
D_SIZE = D_FIN + FIN_SIZE
D_TC = (D_SIZE + 63) // 64 * 64
D_BLOB = D_TC + 4 * TC_IMG3
PLANAR_PAD = 4
# training-mode detector parameters (WM_DT_* in include/wmb200.h)
DT_RB_W1 = 0
DT_RB_B1 = DT_RB_W1 + 3 * 64 * 64
DT_RB_G1 = DT_RB_B1 + 64
DT_RB_BE1 = DT_RB_G1 + 64
DT_RB_W2 = DT_RB_BE1 + 64
DT_RB_B2 = DT_RB_W2 + 3 * 64 * 64
DT_RB_G2 = DT_RB_B2 + 64
DT_RB_BE2 = DT_RB_G2 + 64
DT_RB_SIZE = DT_RB_BE2 + 64
DT_IN_W = 0
DT_IN_B = DT_IN_W + 7 * 64
DT_RB0 = DT_IN_B + 64
DT_HEAD_W = DT_RB0 + 2 * DT_RB_SIZE
DT_HEAD_B = DT_HEAD_W + 32 * 64
DT_SIZE = DT_HEAD_B + 32
DT_STATS = 2 * 4 * 64
# training-mode generator parameters (WM_GT_*)
GT_IN_W = 0
GT_IN_B = GT_IN_W + 7 * 64
GT_RB0 = GT_IN_B + 64
GT_RB1 = GT_RB0 + DT_RB_SIZE
GT_LSTM_WIH = GT_RB1 + DT_RB_SIZE
GT_LSTM_WHH = GT_LSTM_WIH + 256 * 64
GT_LSTM_BIH = GT_LSTM_WHH + 256 * 64
GT_LSTM_BHH = GT_LSTM_BIH + 256
GT_CT_W = GT_LSTM_BHH + 256
GT_CT_B = GT_CT_W + 7 * 64 * 64
GT_RB2 = GT_CT_B + 64
GT_HEAD_W = GT_RB2 + DT_RB_SIZE
GT_HEAD_B = GT_HEAD_W + 64
GT_EMB = GT_HEAD_B + 64
GT_SIZE = GT_EMB + 65536 * 64
GT_STATS = 3 * 4 * 64
MAX_HEAD = 32
POST_FIR, POST_CLAMP, POST_RMS, POST_ALL = 1, 2, 4, 7
MATH_FP32, MATH_BF16X2 = 0, 1

_p = C.c_void_p
_i = C.c_int
_f = C.c_float
_sz = C.c_size_t
_i64 = C.c_int64
_ll = C.c_longlong

# name -> (restype, argtypes); every entry here must be declared in include/wmb200.h
SIGNATURES = {
    "wm_abi_version": (_i, []),
    "wm_last_error": (C.c_char_p, []),
    "wm_device_ok": (_i, []),
    "wm_set_math_mode": (_i, [_i]),
    "wm_get_math_mode": (_i, []),
    "wm_launch_count": (C.c_ulonglong, []),
    "wm_finalize_generator_blob": (_i, [_p, _p]),
    "wm_finalize_detector_blob": (_i, [_p, _p]),
    "wm_planar_bytes": (_sz, [_i, _i]),
    "wm_to_planar": (_i, [_p, _p, _p, _i, _i, _p]),
    "wm_from_planar": (_i, [_p, _p, _i, _i, _p]),
    "wm_conv64_tc_weight_bytes": (_sz, [_i]),
    "wm_pack_conv64_tc": (_i, [_p, _p, _i, _p]),
    "wm_conv64_tc_fwd": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p]),
    "wm_resblock_tc_fwd": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _p]),
    "wm_resblock_tc_hostbias_fwd": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _p]),
    "wm_pack_lstm_tc": (_i, [_p, _p, _p, _p, _p, _p]),
    "wm_debug_lstm_profile": (_i, [_p]),
    "wm_debug_lstm_opts": (_i, [_i]),
    "wm_lstm_tc_fwd": (_i, [_p, _p, _p, _p, _p, _i, _i, _p]),
    "wm_conv_in_k7_fwd": (_i, [_p, _p, _p, _p, _i, _i, _p]),
    "wm_conv64_fwd": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p]),
    "wm_lstm_fwd": (_i, [_p, _p, _p, _p, _p, _i, _i, _p]),
    "wm_head_fwd": (_i, [_p, _p, _p, _p, _i, _i, _i, _p]),
    "wm_postprocess_fwd": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _f, _f, _f, _p]),
    "wm_detect_heads_fwd": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _p]),
    "wm_generator_workspace_bytes": (_sz, [_i, _i]),
    "wm_generator_fwd": (_i, [_p, _p, _i64, _p, _p, _p, _p, _sz, _i, _i, _p]),
    "wm_detector_workspace_bytes": (_sz, [_i, _i]),
    "wm_detector_fwd": (_i, [_p, _p, _p, _p, _sz, _i, _i, _i, _p]),
    "wm_detect_fwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _sz, _i, _i, _i, _p]),
    "wm_embed_detect_workspace_bytes": (_sz, [_i, _i]),
    "wm_embed_detect_fwd": (_i, [_p, _p, _i64, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _sz,
                                 _i, _i, _i, _i, _p]),
    "wm_loss_workspace_bytes": (_sz, [_i, _i]),
    "wm_stft_frames": (_i, [_i, _i]),
    "wm_stft_mag_fwd": (_i, [_p, _p, _i, _i, _i, _i, _p]),
    "wm_hf_penalty_fwd": (_i, [_p, _p, _p, _sz, _i, _i, _i, _i, _p]),
    "wm_loud_fwd": (_i, [_p, _p, _p, _p, _sz, _i, _i, _i, _i, _f, _p]),
    "wm_mel_log_l1_fwd": (_i, [_p, _p, _p, _p, _i, _p, _p, _sz, _i, _i, _i, _i, _p]),
    "wm_bce_heads_fwd": (_i, [_p, _p, _p, _p, _p, _sz, _i, _i, _i, _i, _p]),
    "wm_abs_mean_fwd": (_i, [_p, _p, _p, _sz, _i, _i, _p]),
    "wm_conv1d_out_len": (_i, [_i, _i, _i, _i]),
    "wm_convtranspose1d_out_len": (_i, [_i, _i, _i, _i]),
    "wm_conv1d_fwd": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p]),
    "wm_convtranspose1d_fwd": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p]),
    "wm_lstm_small_fwd": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _p]),
    "wm_convtranspose1d_phase_weight_floats": (_sz, [_i, _i, _i]),
    "wm_convtranspose1d_pack": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _p]),
    "wm_convtranspose1d_phase_fwd": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p]),
    "wm_pconv_plane_rows": (_ll, [_i, _i]),
    "wm_pconv_desc_bytes": (_sz, []),
    "wm_pconv_weight_bytes": (_sz, [_ll, _i]),
    "wm_pconv_pack": (_i, [_p, _p, _ll, _i, _p]),
    "wm_pconv_fwd": (_i, [_p, _p]),
    "wm_pconv_in_fwd": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _ll, _p]),
    "wm_m14_tail8_fwd": (_i, [_p, _ll, _i, _i, _p, _p, _p, _p, _p, _p, _p, _i, _p]),
    "wm_pconv_to_planar": (_i, [_p, _p, _i, _i, _i, _ll, _p]),
    "wm_pconv_from_planar": (_i, [_p, _p, _i, _i, _i, _i, _ll, _p]),
    "wm_resample_fwd": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p]),
    "wm_pcm16_quantize_fwd": (_i, [_p, _p, _sz, _p]),
    "wm_pcm16_dequantize_fwd": (_i, [_p, _p, _sz, _f, _p]),
    "wm_file_metrics_fwd": (_i, [_p, _p, _p, _p, _i, _i, _p]),
    "wm_biquad_workspace_bytes": (_sz, [_i, _ll]),
    "wm_biquad_fwd": (_i, [_p, _p, _p, _i, _ll, _p, _p, _i, _p, _sz, _p]),
    "wm_confusion_counts_fwd": (_i, [_p, _ll, _p, _ll, _f, _p, _p]),
    "wm_roc_points_fwd": (_i, [_p, _ll, _p, _ll, _p, _i, _p, _p, _p]),
    "wm_auc_pairs_fwd": (_i, [_p, _ll, _p, _ll, _p, _p]),
    "wm_stft_bwd_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "wm_hf_penalty_bwd": (_i, [_p, _p, _p, _sz, _i, _i, _i, _i, _f, _i, _p]),
    "wm_loud_bwd": (_i, [_p, _p, _p, _p, _sz, _i, _i, _i, _i, _f, _f, _i, _p]),
    "wm_mel_log_l1_bwd": (_i, [_p, _p, _p, _p, _i, _p, _p, _sz, _i, _i, _i, _i, _f, _i, _p]),
    "wm_abs_mean_bwd": (_i, [_p, _p, _i, _i, _f, _i, _p]),
    "wm_postprocess_bwd": (_i, [_p, _p, _p, _p, _p, _sz, _i, _i, _i, _f, _f, _f, _p]),
    "wm_detector_train_workspace_bytes": (_sz, [_i, _i, _i]),
    "wm_detector_train_step": (_i, [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _f, _f, _f, _f, _f, _f, _i, _p, _p, _p,
                                    _sz, _p]),
    "wm_train_step_workspace_bytes": (_sz, [_i, _i, _i]),
    "wm_train_forward_backward": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _p, _i, _i, _i, _p, _p, _p, _sz,
                                       _p]),
    "wm_bce_heads_bwd": (_i, [_p, _p, _p, _i, _i, _i, _i, _f, _f, _p]),
    "wm_head_bwd_workspace_bytes": (_sz, [_ll, _i]),
    "wm_head_bwd": (_i, [_p, _p, _p, _p, _p, _p, _ll, _i, _p, _sz, _p]),
    "wm_conv_in_k7_bwd_workspace_bytes": (_sz, [_i, _i]),
    "wm_conv_in_k7_bwd": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _p, _sz, _p]),
    "wm_bn_train_workspace_bytes": (_sz, [_ll]),
    "wm_bn_train_fwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _ll, _i, _p, _sz, _p]),
    "wm_bn_train_bwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _ll, _p, _sz, _p]),
    "wm_conv64_bwd_workspace_bytes": (_sz, [_i, _i, _i]),
    "wm_conv64_bwd": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _p, _sz, _p]),
    "wm_conv64_train_fwd": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _p, _sz, _p]),
    "wm_lstm_train_fwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _p]),
    "wm_lstm_train_bwd_workspace_bytes": (_sz, [_i, _i]),
    "wm_lstm_train_bwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _p, _sz, _p]),
    "wm_adam_step": (_i, [_p, _p, _p, _p, _ll, _f, _f, _f, _f, _i, _p]),
    "wm_embed_detect_host_workspace_bytes": (_sz, [_i, _i, _i]),
    "wm_embed_detect_host": (_i, [_p, _p, _i64, _p, _p, _p, _p, _p, _p, _p, _p, _p, _sz,
                                  _i, _i, _i, _i, _i, _p]),
    "wm_embed_detect_host_ragged": (_i, [_p, _p, _i64, _p, _p, _p, _p, _ll, _p, _p, _p, _p, _p, _sz,
                                         _i, _i, _i, _i, _i, _p]),
}


class WmError(RuntimeError):
    pass


_lib = None


def load(path: str | None = None) -> C.CDLL:
    """dlopen libwmb200.so and attach the prototypes.  Raises if it is not built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise WmError(
            f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU or PyTorch fallback for this path)")
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    got = lib.wm_abi_version()
    if got != ABI_VERSION:
        raise WmError(f"libwmb200 ABI {got} != binding ABI {ABI_VERSION}; rebuild the library")
    _lib = lib
    return lib


def last_error() -> str:
    return load().wm_last_error().decode(errors="replace")


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise WmError(f"{what} failed (rc={rc}): {last_error()}")


def ptr(t) -> int | None:
    """Device (or pinned host) address of a torch tensor, None for None."""
    if t is None:
        return None
    return t.data_ptr()
