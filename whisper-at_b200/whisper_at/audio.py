"""Audio front end with the reference's names and semantics
(package/whisper-at/whisper_at/audio.py); the log-mel itself runs in libwat's fused CUDA kernel."""
from __future__ import annotations

import subprocess
from typing import Dict, Optional, Tuple, Union

import numpy as np
import torch
import torch.nn.functional as F

from . import _lib

# audio.py:13-23
SAMPLE_RATE = 16000
N_FFT = 400
N_MELS = 80
HOP_LENGTH = 160
CHUNK_LENGTH = 30
N_SAMPLES = CHUNK_LENGTH * SAMPLE_RATE          # 480000
N_FRAMES = N_SAMPLES // HOP_LENGTH              # 3000
N_SAMPLES_PER_TOKEN = HOP_LENGTH * 2
FRAMES_PER_SECOND = SAMPLE_RATE // HOP_LENGTH
TOKENS_PER_SECOND = SAMPLE_RATE // N_SAMPLES_PER_TOKEN


def load_audio(file: str, sr: int = SAMPLE_RATE) -> np.ndarray:
    """Decode `file` to mono float32 at `sr` Hz through the ffmpeg CLI (audio.py:26-63)."""
    cmd = ["ffmpeg", "-nostdin", "-threads", "0", "-i", file, "-f", "s16le", "-ac", "1", "-acodec", "pcm_s16le",
           "-ar", str(sr), "-"]
    try:
        out = subprocess.run(cmd, capture_output=True, check=True).stdout
    except FileNotFoundError as e:
        raise RuntimeError("Failed to load audio: the ffmpeg executable was not found") from e
    except subprocess.CalledProcessError as e:
        raise RuntimeError(f"Failed to load audio: {e.stderr.decode()}") from e
    return np.frombuffer(out, np.int16).flatten().astype(np.float32) / 32768.0


def pad_or_trim(array, length: int = N_SAMPLES, *, axis: int = -1):
    """Zero-pad or cut `array` to `length` along `axis` (audio.py:66-89); torch tensors and numpy arrays."""
    n = array.shape[axis]
    if torch.is_tensor(array):
        if n > length:
            array = array.narrow(axis, 0, length)
        elif n < length:
            ax = axis % array.ndim
            pads = [0, 0] * array.ndim
            pads[2 * (array.ndim - 1 - ax) + 1] = length - n
            array = F.pad(array, pads)
        return array
    if n > length:
        array = array.take(indices=range(length), axis=axis)
    elif n < length:
        widths = [(0, 0)] * array.ndim
        widths[axis] = (0, length - n)
        array = np.pad(array, widths)
    return array


# one weight-less libwat handle per (device, n_mels) serves log_mel_spectrogram()
_mel_handles: Dict[Tuple[int, int], int] = {}


def _mel_handle(device: torch.device, n_mels: int):
    key = (device.index if device.index is not None else torch.cuda.current_device(), n_mels)
    if key not in _mel_handles:
        import ctypes as C
        L = _lib.lib()
        cfg = _lib.WatConfig(n_mels, 1500, 128, 2, 1, 0, 0, 527, _lib.WAT_FP32, 1)
        h = C.c_void_p()
        with torch.cuda.device(key[0]):
            _lib.check(L.wat_create(C.byref(cfg), C.byref(h)))
        _mel_handles[key] = h
    return _mel_handles[key]


def log_mel_spectrogram(audio: Union[str, np.ndarray, torch.Tensor], n_mels: int = N_MELS, padding: int = 0,
                        device: Optional[Union[str, torch.device]] = None) -> torch.Tensor:
    """Log-mel spectrogram, [n_mels, n_frames] (or [B, n_mels, n_frames] for a 2-D input), same values as
    the reference function (audio.py:110-157): hann(400) STFT hop 160 centre/reflect, |X|^2, slaney mel,
    log10 clamp 1e-10, max(x, x.max() - 8), (x + 4) / 4.  `x.max()` spans the whole input tensor, as in the
    reference.  The reference asserts n_mels == 80; 128 is also accepted here (same slaney generator).

    Runs on the current CUDA device (there is no CPU implementation); the result is returned on the
    device the reference would have returned it on (the input's, or `device`)."""
    if n_mels not in (80, 128):
        raise AssertionError(f"Unsupported n_mels: {n_mels}")
    if not torch.is_tensor(audio):
        if isinstance(audio, str):
            audio = load_audio(audio)
        audio = torch.from_numpy(np.ascontiguousarray(audio))
    if device is not None:
        audio = audio.to(device)
    out_device = audio.device
    if not torch.cuda.is_available():
        raise RuntimeError("whisper_at (B200) needs a CUDA device: log_mel_spectrogram has no CPU path")
    squeeze = audio.ndim == 1
    if audio.ndim not in (1, 2):
        raise ValueError("audio must be 1-D or 2-D")
    x = audio.to(torch.float32)
    if not x.is_cuda:
        x = x.cuda()
    x = x.contiguous()
    if squeeze:
        x = x[None]
    B, n = x.shape
    n_frames = (n + padding) // HOP_LENGTH
    if n_frames < 1 or n <= N_FFT // 2:
        raise RuntimeError("audio too short for a 400-sample reflect-padded STFT")
    L = _lib.lib()
    with torch.cuda.device(x.device):
        h = _mel_handle(x.device, n_mels)
        out = torch.empty((B, n_mels, n_frames), dtype=torch.float32, device=x.device)
        st = torch.cuda.current_stream().cuda_stream
        # scope 1: one clamp floor for the whole tensor (what log_spec.max() does for a batched input)
        _lib.check(L.wat_logmel(h, x.data_ptr(), n, None, n, padding, B, n_frames, 1 if B > 1 else 0,
                                out.data_ptr(), st))
    if squeeze:
        out = out[0]
    return out.to(out_device)
