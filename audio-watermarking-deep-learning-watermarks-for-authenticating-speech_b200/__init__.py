"""wmb200 — B200-native (sm_100a) embed + detect path of the main16 audio watermark.

Drop-in names of the reference (py/main16.py): Generator, Detector, ResBlock,
generate_watermarked_audio, detect_watermark, load_state_dict_strip_prefix, the helper
functions fir_lowpass / clamp_peak / limit_rms and the module constants.  All arithmetic
runs in libwmb200.so (hand-written CUDA, C ABI in include/wmb200.h); there is no CPU or
PyTorch fallback — calls fail loudly when the library or the GPU is missing.
"""
from . import _lib, audio, main14b_2, metrics, ops, packing
from .audio import (Resample, biquad, compute_si_snr, file_metrics, from_pcm16, lowpass_biquad, perceptual_postprocess,
                    resample, save_audio_pcm16, to_pcm16)
from .metrics import auc, classification_report, confusion_counts, roc_curve
from .api import (detect_prob, detect_watermark, evaluate_unseen_file, generate_watermarked_audio, load_audio,
                  process_audio_file_with_delta, run_inference_on_file, save_audio, segment, set_seed)
from .evaluate import evaluate_model, validate_one_epoch
from .train import LR, DetectorTrainer, Trainer
from .training import EarlyStopping, OneCycle, fit, load_ckpt, save_ckpt, train_one_epoch
from .functional import (AUDIO_LEN, HF_PENALTY_W, LAMBDA_DEC, LAMBDA_L1, LAMBDA_LOC, LAMBDA_LOUD,
                         LAMBDA_MSSPEC, MAX_RMS, MESSAGE_BITS, SAMPLE_RATE, bit_targets, clamp_peak,
                         embed_detect, fir_lowpass, limit_rms, postprocess_delta)
from .losses import MultiScaleMelLoss, TFLoudnessLoss, high_freq_penalty, step_losses, stft_magnitude
from .models import Detector, Generator, ResBlock, load_state_dict_strip_prefix
from .sharding import shard_range
from .stream import detect_watermark_folder, embed_detect_stream, process_folder_with_tqdm

__all__ = ["Generator", "Detector", "ResBlock", "generate_watermarked_audio", "detect_watermark", "detect_prob",
           "load_state_dict_strip_prefix", "fir_lowpass", "clamp_peak", "limit_rms", "postprocess_delta",
           "embed_detect", "bit_targets", "shard_range", "load_audio", "save_audio", "segment",
           "SAMPLE_RATE", "AUDIO_LEN", "MESSAGE_BITS", "MAX_RMS", "LAMBDA_L1", "LAMBDA_MSSPEC", "LAMBDA_LOUD",
           "LAMBDA_LOC", "LAMBDA_DEC", "HF_PENALTY_W", "MultiScaleMelLoss", "TFLoudnessLoss", "high_freq_penalty",
           "step_losses", "stft_magnitude", "embed_detect_stream", "process_folder_with_tqdm", "detect_watermark_folder",
           "Resample", "resample", "to_pcm16", "from_pcm16", "file_metrics", "evaluate_model", "validate_one_epoch",
           "DetectorTrainer", "Trainer", "LR", "EarlyStopping", "OneCycle", "fit", "load_ckpt", "save_ckpt",
           "train_one_epoch", "compute_si_snr", "evaluate_unseen_file", "process_audio_file_with_delta",
           "run_inference_on_file", "set_seed", "biquad", "lowpass_biquad", "perceptual_postprocess", "save_audio_pcm16",
           "confusion_counts", "classification_report", "roc_curve", "auc"]
