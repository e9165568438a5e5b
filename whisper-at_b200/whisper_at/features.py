"""Pooled per-layer encoder features in the file format the reference's TL-TR *training* recipe reads
(src/whisper_at_train/dataloader_feat.py:97-125: `np.load(path)['arr_0']`, shape [n_layer, T', d], the loader keeps
or pads to the first 25 pooled steps = 10 s), produced by
src/noise_robust_asr/intermediate_feat_extract/as_full/extract_as_full_whisper_all.py:34-41 with
`np.savez_compressed(path, audio_rep)`.  Here the features come from libwat's encoder (every layer's output,
20x average-pooled)."""
from __future__ import annotations

import math
from typing import Union

import numpy as np
import torch

from .audio import N_SAMPLES, SAMPLE_RATE, log_mel_spectrogram


def pooled_features(model, audio: Union[np.ndarray, torch.Tensor], seconds: float = 10.0) -> np.ndarray:
    """audio: 16 kHz waveform of one clip (<= 30 s).  Returns float32 [n_layer, ceil(seconds * 2.5), d]: the pooled
    states of the clip's first `seconds` (25 steps for the 10 s AudioSet clips the heads are trained on)."""
    a = torch.as_tensor(audio, dtype=torch.float32).flatten()[:N_SAMPLES]
    mel = log_mel_spectrogram(a, n_mels=model.dims.n_mels, padding=N_SAMPLES, device=model.device)[:, :3000]
    all_x = model._encode(mel[None])                       # [L, 75, d]
    steps = min(75, math.ceil(seconds * SAMPLE_RATE / 320 / 20))
    return all_x[:, :steps, :].float().cpu().numpy()


def save_feature_npz(path: str, feat: np.ndarray) -> None:
    """Write one clip's features the way the reference's extraction scripts do (key 'arr_0')."""
    assert feat.ndim == 3
    np.savez_compressed(path, feat.astype(np.float32))


def load_feature_npz(path: str) -> np.ndarray:
    """dataloader_feat.py:97-103."""
    return np.load(path)["arr_0"]
