"""GPU BASELINE for bench.py: the tagging path written in plain PyTorch eager ops on the same B200.

This is "what PyTorch dispatches today" for the reference's modules (the incumbent
Blackwell kernels): torch.stft (cuFFT), F.conv1d (cuDNN), F.linear (cuBLASLt), F.layer_norm, F.gelu and either the
reference's own materialised-qk attention (package/whisper-at/whisper_at/model.py:92-107) or
F.scaled_dot_product_attention (flash / cuDNN fused attention).  Dtype rules follow the reference modules run in half
precision: LayerNorm in fp32 then cast back (model.py:29-31), Linear/Conv weights cast to the activation dtype
(model.py:34-49), softmax in fp32 (model.py:104-106), TL-TR classifier in fp32 (model.py:378), low-compute head entirely
in fp32 (model.py:371).  Two deliberate favours to the baseline: the log-mel is computed on the GPU for the whole batch
(the reference computes it on the CPU, transcribe.py:127) and every clip's pooled states are kept (the reference's
encoder keeps clip 0 only, model.py:174), so one batched call does the same work as our wat_tag.

BENCH INFRASTRUCTURE ONLY: imported by bench.py (`gpu_baseline`, `--impl torch_eager`) and by the GPU tests that check the
baseline itself is a faithful implementation.  Nothing under whisper-at_b200/ imports it, and it drives none of our kernels.
"""
from __future__ import annotations

import math
from typing import Dict

import numpy as np
import torch
import torch.nn.functional as F


def _slaney_filterbank(n_mels: int) -> np.ndarray:
    """librosa.filters.mel(sr=16000, n_fft=400, n_mels) - the generator of the reference's mel_filters.npz (audio.py:92-107)."""
    f_sp, min_log_hz = 200.0 / 3, 1000.0
    min_log_mel, logstep = min_log_hz / f_sp, np.log(6.4) / 27.0
    hz2mel = lambda f: np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-10) / min_log_hz) / logstep, f / f_sp)
    mel2hz = lambda m: np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)
    freqs = np.fft.rfftfreq(400, 1.0 / 16000)
    mel_f = mel2hz(np.linspace(hz2mel(np.float64(0.0)), hz2mel(np.float64(8000.0)), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, freqs)
    w = np.zeros((n_mels, 201), dtype=np.float32)
    for i in range(n_mels):
        w[i] = np.maximum(0, np.minimum(-ramps[i] / fdiff[i], ramps[i + 2] / fdiff[i + 1]))
    w *= (2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels]))[:, None]
    return w


class EagerTagger:
    def __init__(self, sd: Dict[str, torch.Tensor], n_head: int, n_mels: int, device, dtype=torch.bfloat16,
                 attention: str = "sdpa"):
        assert attention in ("sdpa", "materialized")
        self.dev, self.dtype, self.attention, self.n_head, self.n_mels = torch.device(device), dtype, attention, n_head, n_mels
        self.low = "at_model.down_layer.1.weight" in sd
        self.w32 = {k: v.to(self.dev, torch.float32) for k, v in sd.items()}
        self.wh = {k: v.to(dtype) for k, v in self.w32.items() if v.ndim >= 2}          # Linear / Conv weights in the activation dtype
        self.bh = {k: v.to(dtype) for k, v in self.w32.items() if v.ndim == 1}
        self.fb = torch.from_numpy(_slaney_filterbank(n_mels)).to(self.dev)
        self.window = torch.hann_window(400, device=self.dev)
        self.n_layer = 0
        while f"encoder.blocks.{self.n_layer}.attn.query.weight" in sd:
            self.n_layer += 1
        d = sd["encoder.conv1.weight"].shape[0]
        inc = np.log(10000.0) / (d // 2 - 1)
        inv = torch.exp(-inc * torch.arange(d // 2))
        st = torch.arange(1500)[:, None] * inv[None, :]
        self.pos = torch.cat([torch.sin(st), torch.cos(st)], dim=1).to(self.dev, dtype)

    # -------------------------------------------------------------------------------- audio.py:110-157, batched, per-clip max
    def mel(self, audio: torch.Tensor) -> torch.Tensor:
        a = F.pad(audio.to(self.dev, torch.float32), (0, 480000 + 480000 - audio.shape[1]))
        spec = torch.stft(a, 400, 160, window=self.window, return_complex=True)[..., :3000]      # only the frames the clip uses
        mag = spec.abs() ** 2
        ls = torch.clamp(self.fb @ mag, min=1e-10).log10()
        # the clamp floor is the clip's own maximum over its whole padded signal; the padding is silence, so the maximum
        # sits in the first 3001 frames
        ls = torch.maximum(ls, ls.amax(dim=(1, 2), keepdim=True) - 8.0)
        return (ls + 4.0) / 4.0

    # -------------------------------------------------------------------------------- model.py:29-49
    def _ln(self, x, p, dtype=None):
        return F.layer_norm(x.float(), (x.shape[-1],), self.w32[p + ".weight"], self.w32[p + ".bias"], 1e-5).to(dtype or x.dtype)

    def _lin(self, x, p, bias=True):
        if x.dtype == torch.float32:
            return F.linear(x, self.w32[p + ".weight"], self.w32[p + ".bias"] if bias else None)
        return F.linear(x, self.wh[p + ".weight"], self.bh[p + ".bias"] if bias else None)

    def _attn(self, x, p, n_head):                                                       # model.py:70-107
        N, T, D = x.shape
        q, k, v = self._lin(x, p + ".query"), self._lin(x, p + ".key", bias=False), self._lin(x, p + ".value")
        hd = D // n_head
        q = q.view(N, T, n_head, hd).permute(0, 2, 1, 3)
        k = k.view(N, T, n_head, hd).permute(0, 2, 1, 3)
        v = v.view(N, T, n_head, hd).permute(0, 2, 1, 3)
        if self.attention == "sdpa":
            o = F.scaled_dot_product_attention(q, k, v, scale=float(hd) ** -0.5)
        else:
            scale = hd ** -0.25
            qk = ((q * scale) @ (k * scale).transpose(-1, -2)).float()
            o = F.softmax(qk, dim=-1).to(q.dtype) @ v
        return self._lin(o.permute(0, 2, 1, 3).flatten(start_dim=2), p + ".out")

    def _block(self, x, p, n_head):                                                      # model.py:128-139
        x = x + self._attn(self._ln(x, p + ".attn_ln"), p + ".attn", n_head)
        h = F.gelu(self._lin(self._ln(x, p + ".mlp_ln"), p + ".mlp.0"))
        return x + self._lin(h, p + ".mlp.2")

    # -------------------------------------------------------------------------------- model.py:156-177
    def encoder(self, mel: torch.Tensor) -> torch.Tensor:
        x = mel.to(self.dtype)
        x = F.gelu(F.conv1d(x, self.wh["encoder.conv1.weight"], self.bh["encoder.conv1.bias"], padding=1))
        x = F.gelu(F.conv1d(x, self.wh["encoder.conv2.weight"], self.bh["encoder.conv2.bias"], stride=2, padding=1))
        x = (x.permute(0, 2, 1) + self.pos).to(self.dtype)
        pooled = []
        for i in range(self.n_layer):
            x = self._block(x, f"encoder.blocks.{i}", self.n_head)
            pooled.append(x.reshape(x.shape[0], 75, 20, x.shape[-1]).mean(dim=2))         # avg_pool2d (20,1), every clip
        return torch.stack(pooled, dim=1)

    # -------------------------------------------------------------------------------- model.py:351-379
    def head(self, pooled: torch.Tensor, time_resolution: float = 10) -> torch.Tensor:
        B, L, Tp, d = pooled.shape
        dw = int(time_resolution * 2.5)
        S = math.ceil(Tp / dw)
        x = pooled
        if S * dw != Tp:
            x = F.pad(x, (0, 0, 0, S * dw - Tp))
        x = x.reshape(B, L, S, dw, d).permute(0, 2, 1, 3, 4).reshape(B * S * L, dw, d)
        if self.low:
            x = self._lin(self._ln(x.float(), "at_model.down_layer.0"), "at_model.down_layer.1")
        x = self._block(x, "at_model.time_tr", 1)
        x = x.mean(dim=1).reshape(B * S, L, -1)
        x = self._block(x, "at_model.layer_tr", 8)
        x = x.mean(dim=1).float()
        x = self._lin(self._ln(x, "at_model.mlp_layer.0"), "at_model.mlp_layer.1")
        return x.reshape(B, S, -1)

    @torch.no_grad()
    def tag(self, audio: torch.Tensor, time_resolution: float = 10, chunk: int = 0) -> torch.Tensor:
        """audio [B, <=480000] on the device -> logits [B, S, 527] fp32"""
        B = audio.shape[0]
        chunk = chunk or B
        out = []
        for b0 in range(0, B, chunk):
            out.append(self.head(self.encoder(self.mel(audio[b0:b0 + chunk])), time_resolution))
        return torch.cat(out, dim=0)
