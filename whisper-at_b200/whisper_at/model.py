"""`Whisper` model object for the tagging path with the reference's attribute surface
(package/whisper-at/whisper_at/model.py): `.dims`, `.device`, `.is_multilingual`, `.encoder(mel)`,
`.at_model(audio_rep, time_resolution)`, `.embed_audio`, `.transcribe`, `load_state_dict` with the reference's
state-dict keys.  The modules below only hold parameters; the arithmetic runs in libwat (CUDA, sm_100a).

Differences from the reference, all on purpose:
  * `encoder(mel)` returns pooled states for EVERY clip of a batch ([B, L, 75, d]; [L, 75, d] when B == 1);
    the reference keeps only clip 0 (model.py:174).
  * outputs are fp32; the compute precision is the model's `precision` ("bf16": bf16 operands / fp32
    accumulate on tensor cores with an fp32 residual stream; "fp32": everything in fp32).
  * the ASR text decoder is not part of this path: `decoder.*` weights are accepted and ignored.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import numpy as np
import torch
from torch import Tensor, nn

from . import _lib
from .synth import sinusoid_table, tagging_state_shapes


@dataclass
class ModelDimensions:           # model.py:16-27
    n_mels: int
    n_audio_ctx: int
    n_audio_state: int
    n_audio_head: int
    n_audio_layer: int
    n_vocab: int
    n_text_ctx: int
    n_text_state: int
    n_text_head: int
    n_text_layer: int


class _Tensors(nn.Module):
    """weight (+ bias) holder; names match nn.Linear / nn.LayerNorm / nn.Conv1d state-dict keys."""

    def __init__(self, weight_shape, bias: bool = True, norm: bool = False):
        super().__init__()
        w = torch.ones(weight_shape) if norm else torch.zeros(weight_shape)
        self.weight = nn.Parameter(w, requires_grad=False)
        if bias:
            self.bias = nn.Parameter(torch.zeros(weight_shape[0]), requires_grad=False)


class _Attention(nn.Module):     # model.py:61-68
    def __init__(self, d: int):
        super().__init__()
        self.query = _Tensors((d, d))
        self.key = _Tensors((d, d), bias=False)
        self.value = _Tensors((d, d))
        self.out = _Tensors((d, d))


class _Block(nn.Module):         # model.py:110-126 (no cross attention on this path)
    def __init__(self, d: int, n_head: int):
        super().__init__()
        self.n_head = n_head
        self.attn = _Attention(d)
        self.attn_ln = _Tensors((d,), norm=True)
        self.mlp = nn.ModuleDict({"0": _Tensors((4 * d, d)), "2": _Tensors((d, 4 * d))})
        self.mlp_ln = _Tensors((d,), norm=True)


class AudioEncoder(nn.Module):
    """Parameter holder + callable with the reference signature (model.py:142-177)."""

    def __init__(self, n_mels: int, n_ctx: int, n_state: int, n_head: int, n_layer: int):
        super().__init__()
        self.conv1 = _Tensors((n_state, n_mels, 3))
        self.conv2 = _Tensors((n_state, n_state, 3))
        self.register_buffer("positional_embedding", sinusoid_table(n_ctx, n_state))
        self.blocks = nn.ModuleList([_Block(n_state, n_head) for _ in range(n_layer)])
        self.ln_post = _Tensors((n_state,), norm=True)
        self._owner = None

    def forward(self, x: Tensor) -> Tuple[Tensor, Tensor]:
        """x [B, n_mels, 3000] -> (ln_post(x) [B, 1500, d], all_x [L, 75, d] or [B, L, 75, d])."""
        return self._owner[0]._encode(x, want_x=True)


class ATModel(nn.Module):
    """TL-TR head parameters (model.py:321-349); `forward` mirrors model.py:351-379."""

    def __init__(self, label_dim: int = 527, n_layer: int = 32, rep_dim: int = 1280, mode: str = "tl_down_tr_512_1_8"):
        super().__init__()
        self.mode, self.n_layer, self.rep_dim, self.label_dim = mode, n_layer, rep_dim, label_dim
        parts = mode.split("_")
        self.num_tatt_head, self.num_latt_head = int(parts[-2]), int(parts[-1])
        if (self.num_tatt_head, self.num_latt_head) != (1, 8):
            raise ValueError("libwat implements the released TL-TR variants: 1 time head, 8 layer heads")
        if "tl_down_tr" in mode:
            self.inter_rep_dim = int(parts[-3])
            self.down_layer = nn.ModuleDict({"0": _Tensors((rep_dim,), norm=True), "1": _Tensors((self.inter_rep_dim, rep_dim))})
            di = self.inter_rep_dim
        elif "tl_tr" in mode:
            di = rep_dim
        else:
            raise ValueError(f"unknown ATModel mode {mode}")
        self.time_tr = _Block(di, self.num_tatt_head)
        self.layer_tr = _Block(di, self.num_latt_head)
        self.mlp_layer = nn.ModuleDict({"0": _Tensors((di,), norm=True), "1": _Tensors((label_dim, di))})
        self._owner = None

    def forward(self, audio_rep: Tensor, time_resolution=10) -> Tensor:
        return self._owner[0]._head(audio_rep, time_resolution)


class _Engine:
    """One libwat handle (one precision) built from a state dict."""

    def __init__(self, dims: ModelDimensions, low: bool, precision: str, state: Dict[str, Tensor], device: torch.device,
                 max_batch: int):
        if device.type != "cuda":
            raise RuntimeError("whisper_at (B200) runs on CUDA only: there is no CPU implementation of the tagging path")
        self.L = _lib.lib()
        self.device = device
        self.dims = dims
        cfg = _lib.WatConfig(dims.n_mels, dims.n_audio_ctx, dims.n_audio_state, dims.n_audio_head, dims.n_audio_layer,
                             int(low), 512 if low else 0, 527, {"fp32": _lib.WAT_FP32, "bf16": _lib.WAT_BF16}[precision],
                             max_batch)
        self.h = C.c_void_p()
        with torch.cuda.device(device):
            _lib.check(self.L.wat_create(C.byref(cfg), C.byref(self.h)))
            try:
                for key, t in state.items():
                    if key.startswith("decoder."):
                        continue
                    a = t.detach().to("cpu", torch.float32).contiguous()
                    _lib.check(self.L.wat_set_weight(self.h, key.encode(), a.data_ptr(), a.numel()))
                _lib.check(self.L.wat_finalize(self.h))
            except Exception:
                self.L.wat_destroy(self.h)
                self.h = None
                raise

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.L.wat_destroy(self.h)
        except Exception:
            pass

    def launches(self) -> int:
        return int(self.L.wat_kernel_launches(self.h))


class PendingTags:
    """A queued `tag_batch_host_async` call; keeps its host buffers alive until the result has been taken."""

    def __init__(self, eng, ticket: int, audio: Tensor, out: Tensor):
        self._eng, self._ticket, self._audio, self._out = eng, ticket, audio, out

    def result(self) -> Tensor:
        if self._eng is not None:
            with torch.cuda.device(self._eng.device):
                _lib.check(self._eng.L.wat_tag_host_wait(self._eng.h, self._ticket))
            self._eng = self._audio = None
        return self._out


class Whisper(nn.Module):
    """model.py:224-318, tagging path only."""

    def __init__(self, dims: ModelDimensions, at_low_compute: bool = False, precision: str = "bf16", max_batch: int = 128):
        super().__init__()
        if precision not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' or 'fp32'")
        self.dims = dims
        self.at_low_compute = bool(at_low_compute)
        self.precision = precision
        self.max_batch = max_batch
        self.encoder = AudioEncoder(dims.n_mels, dims.n_audio_ctx, dims.n_audio_state, dims.n_audio_head, dims.n_audio_layer)
        mode = "tl_down_tr_512_1_8" if at_low_compute else "tl_tr_1_8"                    # model.py:243-246
        self.at_model = ATModel(n_layer=dims.n_audio_layer, rep_dim=dims.n_audio_state, mode=mode)
        self.encoder._owner = (self,)
        self.at_model._owner = (self,)
        self._engines: Dict[Tuple[str, str], _Engine] = {}

    # ------------------------------------------------------------------ reference surface
    @property
    def device(self) -> torch.device:
        return next(self.parameters()).device

    @property
    def is_multilingual(self) -> bool:
        return self.dims.n_vocab == 51865

    def embed_audio(self, mel: Tensor):
        return self.encoder(mel)

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        """Accepts the merged OpenAI + AT state dict (__init__.py:187-191); `decoder.*` is not on this path."""
        own = {k: v for k, v in state_dict.items() if not k.startswith("decoder.")}
        res = super().load_state_dict(own, strict=strict, **kw)
        self._engines.clear()
        return res

    def set_precision(self, precision: str) -> "Whisper":
        if precision not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' or 'fp32'")
        self.precision = precision
        return self

    # ------------------------------------------------------------------ engine plumbing
    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._engines = {}
        return out

    def engine(self, precision: Optional[str] = None) -> _Engine:
        precision = precision or self.precision
        dev = self.device
        key = (precision, str(dev))
        if key not in self._engines:
            self._engines[key] = _Engine(self.dims, self.at_low_compute, precision, self.state_dict(), dev, self.max_batch)
        return self._engines[key]

    def _encode(self, mel: Tensor, want_x: bool = False, precision: Optional[str] = None):
        d, L = self.dims.n_audio_state, self.dims.n_audio_layer
        squeeze_in = mel.ndim == 2
        if squeeze_in:
            mel = mel[None]
        assert mel.ndim == 3 and tuple(mel.shape[1:]) == (self.dims.n_mels, 2 * self.dims.n_audio_ctx), "incorrect audio shape"
        eng = self.engine(precision)
        m = mel.to(eng.device, torch.float32).contiguous()
        B = m.shape[0]
        with torch.cuda.device(eng.device):
            pooled = torch.empty((B, L, self.dims.n_audio_ctx // 20, d), dtype=torch.float32, device=eng.device)
            x = torch.empty((B, self.dims.n_audio_ctx, d), dtype=torch.float32, device=eng.device) if want_x else None
            st = torch.cuda.current_stream().cuda_stream
            _lib.check(eng.L.wat_encoder(eng.h, m.data_ptr(), B, pooled.data_ptr(), x.data_ptr() if want_x else None, st))
        all_x = pooled[0] if B == 1 else pooled
        return (x, all_x) if want_x else all_x

    def _head(self, audio_rep: Tensor, time_resolution=10, precision: Optional[str] = None) -> Tensor:
        single = audio_rep.ndim == 3
        if single:
            audio_rep = audio_rep[None]
        assert audio_rep.ndim == 4 and audio_rep.shape[1] == self.dims.n_audio_layer and audio_rep.shape[3] == self.dims.n_audio_state
        eng = self.engine(precision)
        dw = int(time_resolution * 2.5)                                                   # model.py:355
        if dw < 1:
            raise ZeroDivisionError("float division by zero")                             # what math.ceil(len / 0) raises
        rep = audio_rep.to(eng.device, torch.float32).contiguous()
        B, _, T, _ = rep.shape
        S = math.ceil(T / dw)
        with torch.cuda.device(eng.device):
            out = torch.empty((B, S, 527), dtype=torch.float32, device=eng.device)
            st = torch.cuda.current_stream().cuda_stream
            _lib.check(eng.L.wat_tltr(eng.h, rep.data_ptr(), B, T, 0, T, dw, out.data_ptr(), st))
        return out[0] if single else out

    # ------------------------------------------------------------------ batched tagging (the throughput path)
    def tag_batch(self, audio: Tensor, at_time_res=10, n_valid: Optional[np.ndarray] = None,
                  precision: Optional[str] = None) -> Tensor:
        """audio [B, n<=480000] fp32 (or int16 PCM) on the model's device -> logits [B, S, 527] on device.
        One fused libwat call: mel -> encoder -> TL-TR (wat_tag / wat_tag_pcm16)."""
        eng = self.engine(precision)
        assert audio.ndim == 2 and audio.shape[1] <= 480000
        i16 = audio.dtype == torch.int16
        a = audio.to(eng.device).contiguous() if i16 else audio.to(eng.device, torch.float32).contiguous()
        B, n = a.shape
        dw = int(at_time_res * 2.5)
        S = math.ceil(75 / dw)
        nv = None
        if n_valid is not None:
            nv_arr = np.ascontiguousarray(n_valid, dtype=np.int32)
            assert nv_arr.shape == (B,)
            nv = nv_arr.ctypes.data_as(C.c_void_p)
        with torch.cuda.device(eng.device):
            out = torch.empty((B, S, 527), dtype=torch.float32, device=eng.device)
            st = torch.cuda.current_stream().cuda_stream
            fn = eng.L.wat_tag_pcm16 if i16 else eng.L.wat_tag
            _lib.check(fn(eng.h, a.data_ptr(), n, nv, n, B, dw, out.data_ptr(), st))
        return out

    def tag_batch_host(self, audio: Tensor, at_time_res=10, out: Optional[Tensor] = None,
                       precision: Optional[str] = None, n_valid: Optional[np.ndarray] = None) -> Tensor:
        """Same through HOST buffers (wat_tag_host): `audio` is a CPU tensor [B, n] (pinned for full
        speed); returns CPU logits.  H2D copy, compute and D2H copy all happen inside the call."""
        return self.tag_batch_host_async(audio, at_time_res, out, precision, n_valid).result()

    def tag_batch_host_async(self, audio: Tensor, at_time_res=10, out: Optional[Tensor] = None,
                             precision: Optional[str] = None, n_valid: Optional[np.ndarray] = None) -> "PendingTags":
        """Queue one host-buffer tagging call (wat_tag_host_submit) and return at once; `.result()` waits
        (wat_tag_host_wait) and returns the CPU logits.  Two calls may be in flight per model: submitting batch t+1
        before asking for the result of batch t moves its PCM over PCIe while batch t computes.  `audio` and `out`
        should be pinned and must not be modified until `.result()` returns."""
        eng = self.engine(precision)
        assert audio.ndim == 2 and audio.shape[1] <= 480000 and not audio.is_cuda
        i16 = audio.dtype == torch.int16                           # 16-bit PCM: half the H2D bytes (wat_tag_host_submit_pcm16)
        a = audio.contiguous() if i16 else audio.to(torch.float32).contiguous()
        B, n = a.shape
        dw = int(at_time_res * 2.5)
        S = math.ceil(75 / dw)
        if out is None:
            out = torch.empty((B, S, 527), dtype=torch.float32)
        assert out.shape == (B, S, 527) and out.dtype == torch.float32 and out.is_contiguous() and not out.is_cuda
        nv = None
        if n_valid is not None:
            nv_arr = np.ascontiguousarray(n_valid, dtype=np.int32)
            assert nv_arr.shape == (B,)
            nv = nv_arr.ctypes.data_as(C.c_void_p)
        ticket = C.c_int64(0)
        with torch.cuda.device(eng.device):
            fn = eng.L.wat_tag_host_submit_pcm16 if i16 else eng.L.wat_tag_host_submit
            _lib.check(fn(eng.h, a.data_ptr(), n, nv, n, B, dw, out.data_ptr(), C.byref(ticket)))
        return PendingTags(eng, ticket.value, a, out)

    def kernel_launches(self) -> int:
        return sum(e.launches() for e in self._engines.values())


def expected_state_keys(dims: ModelDimensions, at_low_compute: bool):
    return list(tagging_state_shapes(dims.n_mels, dims.n_audio_state, dims.n_audio_layer, at_low_compute))
