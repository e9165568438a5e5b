"""`transcribe()` for the tagging path: same signature, `at_time_res` checks, window arithmetic and result
dict as the reference (package/whisper-at/whisper_at/transcribe.py:38-403), without the ASR decoder.

What differs: this package has no text decoder of its own (`text` == "", `segments` == [] unless the caller
supplies one, see `asr_decoder`), and multi-window files advance by a fixed 30 s stride; the reference advances
`seek` by what its ASR decoder's timestamp tokens say (transcribe.py:276-343), which is outside this path.  For
audio of up to 30 s (one window, seek == 0) the `audio_tag` rows are the reference's.  All windows of a file go
through the encoder as ONE batch.

ASR hand-off: the same encoder pass that yields the pooled tagging states also yields
`ln_post(x)` - what the reference's `TextDecoder` cross-attends to (model.py:175, 200-222).  Pass
`asr_decoder=fn` and `fn(audio_features, seeks)` is called once per encoder batch with `audio_features`
[n_windows, 1500, d] fp32 on the model's device and the windows' start frames; it returns one dict per window
(`{"text": str, ...}`), which become `segments`, their texts joined into `text`.  The reference instead re-runs
the encoder inside every `model.decode` call of its temperature fall-back loop (transcribe.py:160-198).
"""
from __future__ import annotations

import math
import warnings
from typing import TYPE_CHECKING, Optional, Tuple, Union

import numpy as np
import torch

from .audio import HOP_LENGTH, N_FRAMES, N_SAMPLES, SAMPLE_RATE, log_mel_spectrogram, pad_or_trim

if TYPE_CHECKING:
    from .model import Whisper


def check_at_time_res(at_time_res) -> float:
    """The reference's validation, message and warning (transcribe.py:131-135)."""
    at_decision_window = at_time_res * 100
    assert at_decision_window % 40 == 0, "Audio tagging resolution at_time_res must be an integer multiple of 0.4 second, e.g., 0.4, 0.8, 1.2, etc, current at_time_res={:.2f}.".format(at_time_res)
    if at_decision_window != 1000:
        warnings.warn("Current at_time_res is {:.2f} second, the audio tagging model is trained with time resolution of 10 seconds. Mismatch time resolution may cause an audio tagging performance drop, but won't impact ASR performance.".format(at_time_res), stacklevel=3)
    return at_decision_window


def transcribe(
    model: "Whisper",
    audio: Union[str, np.ndarray, torch.Tensor],
    *,
    verbose: Optional[bool] = None,
    temperature: Union[float, Tuple[float, ...]] = (0.0, 0.2, 0.4, 0.6, 0.8, 1.0),
    compression_ratio_threshold: Optional[float] = 2.4,
    logprob_threshold: Optional[float] = -1.0,
    no_speech_threshold: Optional[float] = 0.6,
    condition_on_previous_text: bool = True,
    initial_prompt: Optional[str] = None,
    word_timestamps: bool = False,
    prepend_punctuations: str = "\"'“¿([{-",
    append_punctuations: str = "\"'.。,，!！?？:：”)]}、",
    at_time_res=10,
    asr_decoder=None,
    **decode_options,
):
    """Tag an audio file / waveform.  Returns {"text", "segments", "language", "at_time_res", "audio_tag"} with
    `audio_tag` a CPU fp32 tensor [ceil(content_frames / (at_time_res*100)), 527] (transcribe.py:209, 397-403).
    `fp16=False` in decode_options selects the fp32 engine, as it selects fp32 in the reference (transcribe.py:115)."""
    precision = "bf16" if decode_options.get("fp16", True) else "fp32"
    if model.device.type != "cuda":
        raise RuntimeError("whisper_at (B200) runs on CUDA only: move the model with .to('cuda')")

    mel = log_mel_spectrogram(audio, n_mels=model.dims.n_mels, padding=N_SAMPLES, device=model.device)
    content_frames = mel.shape[-1] - N_FRAMES

    at_decision_window = check_at_time_res(at_time_res)
    language = decode_options.get("language", None) or "en"      # no decoder here: language-ID is out of scope

    n_rows = math.ceil(content_frames / at_decision_window)
    all_audio_tags = torch.zeros([n_rows, 527])
    seeks = list(range(0, content_frames, N_FRAMES))
    segments = []
    if seeks:
        segs = torch.stack([pad_or_trim(mel[:, s:s + N_FRAMES], N_FRAMES) for s in seeks])       # transcribe.py:241-244
        chunk = max(1, model.max_batch)
        for c0 in range(0, len(seeks), chunk):
            if asr_decoder is not None:                          # one encoder pass serves tagging AND the caller's text decoder
                x_post, all_x = model._encode(segs[c0:c0 + chunk], want_x=True, precision=precision)
            else:
                all_x = model._encode(segs[c0:c0 + chunk], precision=precision)
            if all_x.ndim == 3:
                all_x = all_x[None]
            if asr_decoder is not None:
                for seek, seg in zip(seeks[c0:c0 + chunk], asr_decoder(x_post, seeks[c0:c0 + chunk])):
                    seg = dict(seg)
                    seg.setdefault("seek", seek)
                    seg.setdefault("start", seek * HOP_LENGTH / SAMPLE_RATE)
                    seg.setdefault("end", min(content_frames, seek + N_FRAMES) * HOP_LENGTH / SAMPLE_RATE)
                    segments.append(seg)
            # windows that share an at_start go through the head as ONE batch (wat_tltr takes [n, L, 75 - at_start, d]);
            # there are at most window / gcd(window, 3000) distinct values
            groups = {}
            for i, seek in enumerate(seeks[c0:c0 + chunk]):
                groups.setdefault(math.floor(seek % at_decision_window / 40), []).append((i, seek))      # transcribe.py:255
            placed = {}
            for at_start, members in groups.items():
                idx = torch.tensor([i for i, _ in members], device=all_x.device)
                tags = model._head(all_x.index_select(0, idx)[:, :, at_start:, :], at_time_res, precision=precision).cpu()
                for (i, _), audio_tag in zip(members, tags):
                    placed[i] = audio_tag
            for i, seek in enumerate(seeks[c0:c0 + chunk]):       # rows are written in window order: a later window overwrites
                audio_tag = placed[i]
                cur_start = math.floor(seek / at_decision_window)
                cur_end = min(all_audio_tags.shape[0], cur_start + audio_tag.shape[0])
                all_audio_tags[cur_start:cur_end, :] = audio_tag[0:cur_end - cur_start, :]       # transcribe.py:261-263
    text = "".join(str(seg.get("text", "")) for seg in segments)
    return dict(text=text, segments=segments, language=language, at_time_res=at_time_res, audio_tag=all_audio_tags.cpu())
