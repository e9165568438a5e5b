"""Deterministic synthetic audio clips and weights for the tagging path.

There is no network for checkpoints or AudioSet audio, so parity and throughput are
measured on synthetic 30 s clips and seeded random weights of the named architecture
(SURVEY.md §8d).  Everything here depends only on (name, seed) so the same tensors can be
regenerated in this container (to drive the reference) and on the GPU box.

State-dict key names follow the reference modules:
  encoder.*  -> package/whisper-at/whisper_at/model.py:142-154 (AudioEncoder.__init__)
  at_model.* -> package/whisper-at/whisper_at/model.py:323-349 (ATModel.__init__)
"""
from __future__ import annotations

import math
import zlib
from collections import OrderedDict
from dataclasses import dataclass
from typing import Dict

import numpy as np
import torch

SAMPLE_RATE = 16000
N_SAMPLES = 480000

#: (n_audio_state, n_audio_head, n_audio_layer) of the OpenAI checkpoints
MODEL_SHAPES = {
    "tiny": (384, 6, 4),
    "base": (512, 8, 6),
    "small": (768, 12, 12),
    "medium": (1024, 16, 24),
    "large-v2": (1280, 20, 32),
}


@dataclass(frozen=True)
class TagConfig:
    """One BASELINE.json configuration restated as numbers."""
    name: str
    n_mels: int
    low: bool
    batch: int
    at_time_res: float


BASELINE_CONFIGS = (
    TagConfig("tiny", 80, False, 1, 10),
    TagConfig("base", 80, False, 64, 10),
    TagConfig("small", 80, True, 256, 2),
    TagConfig("medium", 80, True, 512, 10),
    TagConfig("large-v2", 128, False, 1024, 10),
)


def synth_clip(index: int, n_samples: int = N_SAMPLES) -> torch.Tensor:
    """Clip `index` of the synthetic set: fp32 [n_samples] in roughly [-1, 1].

    clip 0 is white noise * 0.1; the others are three AM-modulated sinusoids plus noise
    with a random overall gain; every 8th clip has its last 10 s silenced so the
    `max - 8` clamp of the log-mel is exercised.
    """
    g = torch.Generator().manual_seed(1000 + index)
    if index == 0:
        return torch.randn(n_samples, generator=g) * 0.1
    t = torch.arange(n_samples, dtype=torch.float64) / SAMPLE_RATE
    u = torch.rand(16, generator=g, dtype=torch.float64)
    x = torch.zeros(n_samples, dtype=torch.float64)
    for k in range(3):
        freq = 50.0 + u[k] * (7000.0 - 50.0)
        amp = 0.05 + u[3 + k] * 0.25
        am = 0.5 + u[6 + k] * 3.5
        phase = u[9 + k] * 2 * math.pi
        x += amp * torch.sin(2 * math.pi * freq * t + phase) * (0.5 + 0.5 * torch.sin(2 * math.pi * am * t))
    sigma = 10.0 ** (-3.0 + 2.0 * u[12])
    x += sigma * torch.randn(n_samples, generator=g, dtype=torch.float64)
    x *= 10.0 ** (-2.0 + 2.0 * u[13])
    if index % 8 == 7:
        x[n_samples - 10 * SAMPLE_RATE:] = 0.0
    return x.to(torch.float32)


def synth_batch(batch: int, start: int = 0) -> torch.Tensor:
    return torch.stack([synth_clip(start + i) for i in range(batch)])


def _seed_for(name: str, seed: int) -> int:
    return (zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF


def _block_shapes(prefix: str, d: int) -> "OrderedDict[str, tuple]":
    s = OrderedDict()
    s[f"{prefix}.attn.query.weight"] = (d, d)
    s[f"{prefix}.attn.query.bias"] = (d,)
    s[f"{prefix}.attn.key.weight"] = (d, d)
    s[f"{prefix}.attn.value.weight"] = (d, d)
    s[f"{prefix}.attn.value.bias"] = (d,)
    s[f"{prefix}.attn.out.weight"] = (d, d)
    s[f"{prefix}.attn.out.bias"] = (d,)
    s[f"{prefix}.attn_ln.weight"] = (d,)
    s[f"{prefix}.attn_ln.bias"] = (d,)
    s[f"{prefix}.mlp.0.weight"] = (4 * d, d)
    s[f"{prefix}.mlp.0.bias"] = (4 * d,)
    s[f"{prefix}.mlp.2.weight"] = (d, 4 * d)
    s[f"{prefix}.mlp.2.bias"] = (d,)
    s[f"{prefix}.mlp_ln.weight"] = (d,)
    s[f"{prefix}.mlp_ln.bias"] = (d,)
    return s


def tagging_state_shapes(n_mels: int, d: int, n_layer: int, low: bool, at_dim: int = 512,
                         n_class: int = 527) -> "OrderedDict[str, tuple]":
    """Names and shapes of every tensor the tagging path reads (encoder + at_model)."""
    s = OrderedDict()
    s["encoder.conv1.weight"] = (d, n_mels, 3)
    s["encoder.conv1.bias"] = (d,)
    s["encoder.conv2.weight"] = (d, d, 3)
    s["encoder.conv2.bias"] = (d,)
    for i in range(n_layer):
        s.update(_block_shapes(f"encoder.blocks.{i}", d))
    s["encoder.ln_post.weight"] = (d,)
    s["encoder.ln_post.bias"] = (d,)
    di = at_dim if low else d
    if low:
        s["at_model.down_layer.0.weight"] = (d,)
        s["at_model.down_layer.0.bias"] = (d,)
        s["at_model.down_layer.1.weight"] = (di, d)
        s["at_model.down_layer.1.bias"] = (di,)
    s.update(_block_shapes("at_model.time_tr", di))
    s.update(_block_shapes("at_model.layer_tr", di))
    s["at_model.mlp_layer.0.weight"] = (di,)
    s["at_model.mlp_layer.0.bias"] = (di,)
    s["at_model.mlp_layer.1.weight"] = (n_class, di)
    s["at_model.mlp_layer.1.bias"] = (n_class,)
    return s


def _is_norm(name: str) -> bool:
    return ("_ln." in name or "ln_post" in name or "down_layer.0." in name or "mlp_layer.0." in name)


def synth_state_dict(n_mels: int, d: int, n_layer: int, low: bool, seed: int = 0,
                     init: str = "default") -> Dict[str, torch.Tensor]:
    """Seeded weights. `default` mirrors PyTorch's Linear/Conv/LayerNorm defaults
    (U(-1/sqrt(fan_in), 1/sqrt(fan_in)); LN 1/0). `lively` perturbs LN gains and all biases
    so the network is input-sensitive (last bullet)."""
    assert init in ("default", "lively")
    out: Dict[str, torch.Tensor] = OrderedDict()
    for name, shape in tagging_state_shapes(n_mels, d, n_layer, low).items():
        g = torch.Generator().manual_seed(_seed_for(name, seed))
        if _is_norm(name):
            if init == "default":
                w = torch.ones(shape) if name.endswith("weight") else torch.zeros(shape)
            elif name.endswith("weight"):
                w = 0.5 + torch.rand(shape, generator=g)
            else:
                w = 0.1 * torch.randn(shape, generator=g)
        else:
            fan_in = 1
            if name.endswith("weight"):
                for k in shape[1:]:
                    fan_in *= k
            else:  # bias: fan_in of the matching weight
                wname = name[:-4] + "weight"
                wshape = tagging_state_shapes(n_mels, d, n_layer, low).get(wname)
                if wshape is not None:
                    for k in wshape[1:]:
                        fan_in *= k
            bound = 1.0 / math.sqrt(fan_in)
            if init == "lively" and name.endswith("bias"):
                w = 0.1 * torch.randn(shape, generator=g)
            else:
                w = (torch.rand(shape, generator=g) * 2 - 1) * bound
        out[name] = w.to(torch.float32).contiguous()
    return out


def decoder_state_shapes(d: int, n_layer: int, n_vocab: int, n_ctx: int) -> "OrderedDict[str, tuple]":
    """Parameters of the reference's TextDecoder (model.py:180-198) for the ASR hand-off test."""
    s = OrderedDict()
    s["decoder.token_embedding.weight"] = (n_vocab, d)
    s["decoder.positional_embedding"] = (n_ctx, d)
    for i in range(n_layer):
        p = f"decoder.blocks.{i}"
        s.update(_block_shapes(p, d))
        for k, shape in _block_shapes(p, d).items():
            if ".attn." in k or ".attn_ln." in k:
                s[k.replace(".attn.", ".cross_attn.").replace(".attn_ln.", ".cross_attn_ln.")] = shape
    s["decoder.ln.weight"] = (d,)
    s["decoder.ln.bias"] = (d,)
    return s


def synth_decoder_state_dict(d: int, n_layer: int, n_vocab: int = 1024, n_ctx: int = 64, seed: int = 1) -> Dict[str, torch.Tensor]:
    """Seeded 'lively' weights of a small TextDecoder: the consumer of `ln_post(x)` in the hand-off test."""
    out: Dict[str, torch.Tensor] = OrderedDict()
    for name, shape in decoder_state_shapes(d, n_layer, n_vocab, n_ctx).items():
        g = torch.Generator().manual_seed(_seed_for(name, seed))
        if "_ln." in name or name.startswith("decoder.ln."):
            w = 0.5 + torch.rand(shape, generator=g) if name.endswith("weight") else 0.1 * torch.randn(shape, generator=g)
        elif name.endswith("embedding.weight") or name.endswith("positional_embedding"):
            w = 0.3 * torch.randn(shape, generator=g)
        elif name.endswith("bias"):
            w = 0.1 * torch.randn(shape, generator=g)
        else:
            w = (torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(shape[1])
        out[name] = w.to(torch.float32).contiguous()
    return out


def tltr_state_shapes(mode: str, n_layer: int, rep_dim: int, label_dim: int = 527) -> "OrderedDict[str, tuple]":
    """Parameters of the training recipe's TLTR module for a mode string (src/whisper_at_train/models.py:49-106),
    under the module's own key names."""
    parts = mode.split("_")
    down = "down" in parts
    di = int(parts[parts.index("tr") + 1]) if down else rep_dim
    s = OrderedDict()
    if down:
        s["down_layer.0.weight"] = (rep_dim,)
        s["down_layer.0.bias"] = (rep_dim,)
        s["down_layer.1.weight"] = (di, rep_dim)
        s["down_layer.1.bias"] = (di,)
    if parts[0] == "wa":
        s["layer_weight"] = (n_layer,)
    if "tr" in parts:
        s.update(_block_shapes("time_tr", di))
    if parts[0] == "lw":
        s.update(_block_shapes("layer_tr", di))
    s["mlp_layer.0.weight"] = (di,)
    s["mlp_layer.0.bias"] = (di,)
    s["mlp_layer.1.weight"] = (label_dim, di)
    s["mlp_layer.1.bias"] = (label_dim,)
    return s


def synth_tltr_state_dict(mode: str, n_layer: int, rep_dim: int, label_dim: int = 527, seed: int = 1) -> Dict[str, torch.Tensor]:
    """Seeded 'lively' weights for a TLTR head variant (LN gains U(0.5,1.5), biases N(0,0.1), positive layer weights)."""
    shapes = tltr_state_shapes(mode, n_layer, rep_dim, label_dim)
    out: Dict[str, torch.Tensor] = OrderedDict()
    for name, shape in shapes.items():
        g = torch.Generator().manual_seed(_seed_for("tltr." + mode + "." + name, seed))
        if name == "layer_weight":
            w = 0.2 + torch.rand(shape, generator=g)
        elif _is_norm(name):
            w = 0.5 + torch.rand(shape, generator=g) if name.endswith("weight") else 0.1 * torch.randn(shape, generator=g)
        elif name.endswith("bias"):
            w = 0.1 * torch.randn(shape, generator=g)
        else:
            fan_in = 1
            for k in shape[1:]:
                fan_in *= k
            w = (torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(fan_in)
        out[name] = w.to(torch.float32).contiguous()
    return out


def synth_audio_rep(batch: int, n_layer: int, t: int, rep_dim: int, seed: int = 7) -> torch.Tensor:
    """Pooled-encoder-state stand-in [B, L, T', d] for head-only tests: layer-dependent offset and scale."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(batch, n_layer, t, rep_dim, generator=g)
    scale = 0.5 + torch.arange(n_layer, dtype=torch.float32)[None, :, None, None] / n_layer
    return (x * scale + 0.1 * torch.randn(1, n_layer, 1, rep_dim, generator=g)).contiguous()


def sinusoid_table(length: int, channels: int, max_timescale: float = 10000.0) -> torch.Tensor:
    """Positional table of the encoder (reference model.py:52-58): cat(sin, cos) of
    t * exp(-ln(max_timescale)/(channels/2-1) * i). Computed with the same fp32 torch ops
    so the buffer is bit-identical to the reference's registered buffer."""
    assert channels % 2 == 0
    inc = np.log(max_timescale) / (channels // 2 - 1)
    inv = torch.exp(-inc * torch.arange(channels // 2))
    st = torch.arange(length)[:, None] * inv[None, :]
    return torch.cat([torch.sin(st), torch.cos(st)], dim=1)
