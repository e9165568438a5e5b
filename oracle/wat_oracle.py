"""CPU ORACLE for the Whisper-AT tagging path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this file.  The product (whisper-at_b200/) never does; it fails loudly without its
CUDA library.

What it is: a restatement, in plain torch-CPU / numpy arithmetic, of the reference's
algorithm for  log_mel_spectrogram -> AudioEncoder.forward (all layers' pooled states)
-> ATModel.forward, with *per-clip* semantics for a batch (the reference itself only keeps
clip 0 of a batch, package/whisper-at/whisper_at/model.py:174).

The arithmetic of the reference lives in PyTorch (third-party; unpinned `torch` in
package/whisper-at/requirements.txt:3, torch 2.7.0 in poetry.lock:2654; this image has
2.11.0).  The published algorithms restated here: STFT (framed real DFT, centre/reflect),
slaney mel filterbank (librosa.filters.mel, cited at audio.py:96-101), conv1d, LayerNorm,
scaled dot-product attention with softmax, exact-erf GELU, average pooling.

Parity pinning: the reference's own tests hold no vector for this path, so
the pin is the reference ITSELF run in the build container: oracle/make_golden.py imports
/root/reference, checks every function here against it (bit-exact or <=2e-6) and commits
the reference's outputs as tests/golden/*.npz; tests/test_oracle_cpu.py re-checks this file
against those fixtures wherever it runs.  The head variants of the training recipe
(tltr_variant; src/whisper_at_train/models.py:108-200) are pinned the same way by
oracle/make_golden_tltr.py against the reference's own TLTR class.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

SAMPLE_RATE = 16000
N_FFT = 400
HOP = 160
N_SAMPLES = 480000
N_FRAMES = 3000
N_CTX = 1500
POOL = 20


# --------------------------------------------------------------------------- mel front end
def _hz_to_mel(f):
    f = np.asanyarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-10) / min_log_hz) / logstep, f / f_sp)


def _mel_to_hz(m):
    m = np.asanyarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_filterbank(n_mels: int = 80) -> np.ndarray:
    """Slaney-normalised triangular filterbank, fp32 [n_mels, 201].

    Restates librosa.filters.mel(sr=16000, n_fft=400, n_mels=n) which produced the
    reference's assets/mel_filters.npz (audio.py:92-107).  For n_mels=80 this is
    bit-identical to the npz (checked by make_golden.py); the reference asserts
    n_mels == 80 (audio.py:103), so 128 bins use the same generator."""
    fftfreqs = np.fft.rfftfreq(n=N_FFT, d=1.0 / SAMPLE_RATE)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(0.0), _hz_to_mel(SAMPLE_RATE / 2), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    w = np.zeros((n_mels, 1 + N_FFT // 2), dtype=np.float32)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        w[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    w *= enorm[:, None]
    return w


def stft_power(audio: torch.Tensor, dtype=torch.float32, explicit_dft: bool = False) -> torch.Tensor:
    """|STFT|^2 with n_fft=400, hop=160, periodic hann, center=True/reflect, last frame dropped
    (audio.py:147-149).  Returns [201, n_frames]."""
    x = audio.to(dtype)
    pad = N_FFT // 2
    xp = torch.cat([x[1:pad + 1].flip(0), x, x[-pad - 1:-1].flip(0)])
    frames = xp.unfold(0, N_FFT, HOP)                       # [n_frames+1, 400]
    n = torch.arange(N_FFT, dtype=torch.float64)
    window = (0.5 - 0.5 * torch.cos(2 * math.pi * n / N_FFT)).to(dtype)
    if dtype == torch.float32:
        window = torch.hann_window(N_FFT)                    # the reference's own fp32 window
    fw = frames * window
    if explicit_dft:
        k = torch.arange(N_FFT // 2 + 1, dtype=torch.float64)
        ang = 2 * math.pi * torch.outer(n, k) / N_FFT
        re = fw.to(torch.float64) @ torch.cos(ang)
        im = fw.to(torch.float64) @ (-torch.sin(ang))
        p = (re * re + im * im).to(dtype)
    else:
        spec = torch.fft.rfft(fw, dim=-1)
        p = spec.abs() ** 2
    return p[:-1].transpose(0, 1).contiguous()


def log_mel(audio: torch.Tensor, n_mels: int = 80, padding: int = 0, dtype=torch.float32,
            explicit_dft: bool = False) -> torch.Tensor:
    """log_mel_spectrogram (audio.py:110-157) for ONE clip -> [n_mels, n_frames]."""
    if padding > 0:
        audio = F.pad(audio, (0, padding))
    p = stft_power(audio, dtype, explicit_dft)
    fb = torch.from_numpy(mel_filterbank(n_mels)).to(dtype)
    mel = fb @ p
    ls = torch.clamp(mel, min=1e-10).log10()
    ls = torch.maximum(ls, ls.max() - 8.0)
    return (ls + 4.0) / 4.0


def log_mel_clip(clip: torch.Tensor, n_mels: int = 80, dtype=torch.float32, explicit_dft=False) -> torch.Tensor:
    """What transcribe feeds the encoder for a <=30 s clip: pad 30 s of zeros, keep the
    first 3000 frames (transcribe.py:127, 241-244)."""
    clip = clip[:N_SAMPLES]
    return log_mel(clip, n_mels, padding=N_SAMPLES, dtype=dtype, explicit_dft=explicit_dft)[:, :N_FRAMES]


# --------------------------------------------------------------------------- encoder
def sinusoids(length: int, channels: int, max_timescale: float = 10000.0) -> torch.Tensor:
    """model.py:52-58."""
    inc = np.log(max_timescale) / (channels // 2 - 1)
    inv = torch.exp(-inc * torch.arange(channels // 2))
    st = torch.arange(length)[:, None] * inv[None, :]
    return torch.cat([torch.sin(st), torch.cos(st)], dim=1)


def _gelu(x):
    return 0.5 * x * (1.0 + torch.erf(x * (1.0 / math.sqrt(2.0))))


def _ln(x, w, b):
    return F.layer_norm(x, (x.shape[-1],), w, b, 1e-5)


def _block(x: torch.Tensor, sd: Dict[str, torch.Tensor], p: str, n_head: int) -> torch.Tensor:
    """ResidualAttentionBlock.forward (model.py:128-139) + MultiHeadAttention (model.py:70-107).
    x: [N, T, D]."""
    g = lambda k: sd[f"{p}.{k}"].to(x.dtype)
    N, T, D = x.shape
    h = _ln(x, g("attn_ln.weight"), g("attn_ln.bias"))
    q = h @ g("attn.query.weight").T + g("attn.query.bias")
    k = h @ g("attn.key.weight").T                                   # key has no bias (model.py:66)
    v = h @ g("attn.value.weight").T + g("attn.value.bias")
    hd = D // n_head
    scale = hd ** -0.25
    q = q.view(N, T, n_head, hd).permute(0, 2, 1, 3) * scale
    k = k.view(N, T, n_head, hd).permute(0, 2, 3, 1) * scale
    v = v.view(N, T, n_head, hd).permute(0, 2, 1, 3)
    w = torch.softmax((q @ k).float() if x.dtype != torch.float64 else (q @ k), dim=-1).to(x.dtype)
    a = (w @ v).permute(0, 2, 1, 3).reshape(N, T, D)
    x = x + (a @ g("attn.out.weight").T + g("attn.out.bias"))
    h = _ln(x, g("mlp_ln.weight"), g("mlp_ln.bias"))
    h = _gelu(h @ g("mlp.0.weight").T + g("mlp.0.bias"))
    return x + (h @ g("mlp.2.weight").T + g("mlp.2.bias"))


def encoder_pooled(mel: torch.Tensor, sd: Dict[str, torch.Tensor], n_head: int, dtype=torch.float32,
                   return_x: bool = False):
    """AudioEncoder.forward (model.py:156-177) with per-clip pooled states.
    mel [B, n_mels, 3000] -> pooled [B, L, 75, d] (each clip == the reference run on it alone)."""
    g = lambda k: sd[k].to(dtype)
    x = mel.to(dtype)
    x = _gelu(F.conv1d(x, g("encoder.conv1.weight"), g("encoder.conv1.bias"), padding=1))
    x = _gelu(F.conv1d(x, g("encoder.conv2.weight"), g("encoder.conv2.bias"), stride=2, padding=1))
    x = x.permute(0, 2, 1)
    d = x.shape[-1]
    assert x.shape[1] == N_CTX, "incorrect audio shape"
    x = x + sinusoids(N_CTX, d).to(dtype)
    n_layer = 0
    while f"encoder.blocks.{n_layer}.attn.query.weight" in sd:
        n_layer += 1
    pooled = []
    for i in range(n_layer):
        x = _block(x, sd, f"encoder.blocks.{i}", n_head)
        pooled.append(x.reshape(x.shape[0], N_CTX // POOL, POOL, d).mean(dim=2))
    out = torch.stack(pooled, dim=1)
    if return_x:
        return out, _ln(x, g("encoder.ln_post.weight"), g("encoder.ln_post.bias"))
    return out


# --------------------------------------------------------------------------- ASR hand-off consumer
def _mha(xq, xkv, sd, p, n_head, mask=None):
    """MultiHeadAttention.forward + qkv_attention (model.py:70-107): self- or cross-attention."""
    g = lambda k: sd[f"{p}.{k}"].to(xq.dtype)
    N, Tq, D = xq.shape
    q = xq @ g("query.weight").T + g("query.bias")
    k = xkv @ g("key.weight").T
    v = xkv @ g("value.weight").T + g("value.bias")
    hd = D // n_head
    scale = hd ** -0.25
    q = q.view(N, Tq, n_head, hd).permute(0, 2, 1, 3) * scale
    k = k.view(N, -1, n_head, hd).permute(0, 2, 3, 1) * scale
    v = v.view(N, -1, n_head, hd).permute(0, 2, 1, 3)
    qk = q @ k
    if mask is not None:
        qk = qk + mask[:Tq, :Tq]
    w = torch.softmax(qk.float(), dim=-1).to(xq.dtype)
    return (w @ v).permute(0, 2, 1, 3).reshape(N, Tq, D) @ g("out.weight").T + g("out.bias")


def text_decoder_logits(tokens: torch.Tensor, xa: torch.Tensor, sd: Dict[str, torch.Tensor], n_head: int,
                        dtype=torch.float32) -> torch.Tensor:
    """TextDecoder.forward without a kv cache (model.py:200-222): tokens [B, n] int64, xa [B, 1500, d] = ln_post(x) of
    the audio encoder -> logits [B, n, n_vocab].  The consumer of wat_encoder's x_out in the hand-off test; the decoder
    itself stays PyTorch."""
    g = lambda k: sd[f"decoder.{k}"].to(dtype)
    n = tokens.shape[-1]
    x = g("token_embedding.weight")[tokens] + g("positional_embedding")[:n]
    xa = xa.to(dtype)
    mask = torch.full((n, n), float("-inf")).triu_(1).to(x.device)
    i = 0
    while f"decoder.blocks.{i}.attn.query.weight" in sd:
        p = f"decoder.blocks.{i}"
        h = _ln(x, g(f"blocks.{i}.attn_ln.weight"), g(f"blocks.{i}.attn_ln.bias"))
        x = x + _mha(h, h, sd, p + ".attn", n_head, mask)
        h = _ln(x, g(f"blocks.{i}.cross_attn_ln.weight"), g(f"blocks.{i}.cross_attn_ln.bias"))
        x = x + _mha(h, xa, sd, p + ".cross_attn", n_head)
        h = _ln(x, g(f"blocks.{i}.mlp_ln.weight"), g(f"blocks.{i}.mlp_ln.bias"))
        h = _gelu(h @ g(f"blocks.{i}.mlp.0.weight").T + g(f"blocks.{i}.mlp.0.bias"))
        x = x + (h @ g(f"blocks.{i}.mlp.2.weight").T + g(f"blocks.{i}.mlp.2.bias"))
        i += 1
    x = _ln(x, g("ln.weight"), g("ln.bias"))
    return (x @ g("token_embedding.weight").T).float()


# --------------------------------------------------------------------------- TL-TR head
def decision_window(time_resolution: float) -> int:
    """model.py:355."""
    return int(time_resolution * 2.5)


def tltr_head(pooled: torch.Tensor, sd: Dict[str, torch.Tensor], time_resolution: float = 10,
              dtype=torch.float32, n_time_head: int = 1, n_layer_head: int = 8) -> torch.Tensor:
    """ATModel.forward (model.py:351-379), batched over clips.
    pooled [B, L, T', d] -> logits [B, S, n_class]."""
    low = "at_model.down_layer.1.weight" in sd
    g = lambda k: sd[f"at_model.{k}"].to(dtype)
    B, L, Tp, d = pooled.shape
    dw = decision_window(time_resolution)
    S = math.ceil(Tp / dw)
    x = pooled.to(dtype)
    if S * dw != Tp:
        x = F.pad(x, (0, 0, 0, S * dw - Tp))
    x = x.reshape(B, L, S, dw, d).permute(0, 2, 1, 3, 4).reshape(B * S * L, dw, d)
    if low:
        x = _ln(x, g("down_layer.0.weight"), g("down_layer.0.bias")) @ g("down_layer.1.weight").T + g("down_layer.1.bias")
    x = _block(x, {k: v for k, v in sd.items()}, "at_model.time_tr", n_time_head)
    x = x.mean(dim=1).reshape(B * S, L, -1)
    x = _block(x, sd, "at_model.layer_tr", n_layer_head)
    x = x.mean(dim=1)
    x = _ln(x, g("mlp_layer.0.weight"), g("mlp_layer.0.bias")) @ g("mlp_layer.1.weight").T + g("mlp_layer.1.bias")
    return x.reshape(B, S, -1)


def tltr_variant(audio_rep: torch.Tensor, sd: Dict[str, torch.Tensor], mode: str, dtype=torch.float32) -> torch.Tensor:
    """TLTR.forward of the training recipe (src/whisper_at_train/models.py:108-200) for every mode string it knows.
    audio_rep [B, L, T', d]; sd uses the module's own keys ('time_tr.*', 'mlp_layer.*', 'down_layer.*', 'layer_weight').
    Returns logits [B, label_dim]."""
    g = lambda k: sd[k].to(dtype)
    x = audio_rep.to(dtype)
    B, L = x.shape[0], x.shape[1]
    mlp = lambda v: _ln(v, g("mlp_layer.0.weight"), g("mlp_layer.0.bias")) @ g("mlp_layer.1.weight").T + g("mlp_layer.1.bias")
    down = lambda v: _ln(v, g("down_layer.0.weight"), g("down_layer.0.bias")) @ g("down_layer.1.weight").T + g("down_layer.1.bias")
    wa = lambda v: (v.permute(0, 2, 3, 1) @ g("layer_weight")) / g("layer_weight").sum()          # models.py:152-153
    parts = mode.split("_")
    if mode == "mean_mlp":                                                                       # :113-117
        return mlp(x.mean(dim=1).mean(dim=1))
    if mode == "last_mlp":                                                                       # :120-124
        return mlp(x[:, -1].mean(dim=1))
    if mode == "wa_mlp":                                                                         # :127-132
        v = x.mean(dim=2).permute(0, 2, 1)
        return mlp((v @ g("layer_weight")) / g("layer_weight").sum())
    if mode.startswith("mean_tr"):                                                               # :135-140
        return mlp(_block(x.mean(dim=1), sd, "time_tr", int(parts[-1])).mean(dim=1))
    if mode.startswith("last_tr"):                                                               # :143-148
        return mlp(_block(x[:, -1], sd, "time_tr", int(parts[-1])).mean(dim=1))
    if mode.startswith("wa_tr"):                                                                 # :151-157
        return mlp(_block(wa(x), sd, "time_tr", int(parts[-1])).mean(dim=1))
    if mode.startswith("wa_down_tr"):                                                            # :160-167
        return mlp(_block(down(wa(x)), sd, "time_tr", int(parts[-1])).mean(dim=1))
    if mode.startswith("lw_tr") or mode.startswith("lw_down_tr"):                                # :170-200
        if mode.startswith("lw_down_tr"):
            x = down(x)
        v = _block(x.reshape(B * L, x.shape[2], x.shape[3]), sd, "time_tr", int(parts[-2])).mean(dim=1)
        v = _block(v.reshape(B, L, -1), sd, "layer_tr", int(parts[-1])).mean(dim=1)
        return mlp(v)
    raise ValueError(f"unknown TLTR mode {mode!r}")


# --------------------------------------------------------------------------- whole path
def tag(audio: torch.Tensor, sd: Dict[str, torch.Tensor], n_head: int, n_mels: int = 80,
        time_resolution: float = 10, dtype=torch.float32, at_start: int = 0) -> torch.Tensor:
    """audio [B, <=480000] -> logits [B, S, 527]: the single-window path of transcribe
    (transcribe.py:127, 241-263)."""
    mel = torch.stack([log_mel_clip(a, n_mels) for a in audio])
    pooled = encoder_pooled(mel, sd, n_head, dtype)
    return tltr_head(pooled[:, :, at_start:, :], sd, time_resolution, dtype)


def top_k_labels(logits_row: torch.Tensor, k: int = 5):
    v, i = torch.topk(logits_row, k)
    return i.tolist(), v.tolist()
