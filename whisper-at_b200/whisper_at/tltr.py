"""The TL-TR head family of the training recipe on libwat (reference: src/whisper_at_train/models.py:49-200).

`TLTR(label_dim, n_layer, rep_dim, mode)` has the reference class's constructor, mode strings and state_dict keys
('time_tr.attn.query.weight', 'mlp_layer.1.bias', 'layer_weight', ...; a DataParallel 'module.' prefix is accepted):

    mean_mlp | last_mlp | wa_mlp | mean_tr_{h} | last_tr_{h} | wa_tr_{h} | wa_down_tr_{dim}_{h}          (the paper's baselines)
    lw_tr_{t}_{l} | lw_down_tr_{dim}_{t}_{l}                                                    (TL-TR; the package's heads)

`forward(audio_rep[B, n_layer, T', rep_dim])` returns logits `[B, label_dim]` (fp32, on the input's device).  Inference
only: training stays with the reference's PyTorch module, which these weights come from.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Tuple

import torch

from . import _lib


def parse_mode(mode: str, rep_dim: int) -> Tuple[int, int, int, int]:
    """mode string -> (wat_head_mode, inter_dim, n_time_head, n_layer_head), following models.py:56-106."""
    parts = mode.split("_")
    for name in ("lw_down_tr", "wa_down_tr", "mean_mlp", "last_mlp", "wa_mlp", "mean_tr", "last_tr", "wa_tr", "lw_tr"):
        if mode == name or mode.startswith(name + "_"):
            break
    else:
        raise ValueError(f"unknown TLTR mode {mode!r}")
    code = _lib.HEAD_MODES[name]
    try:
        if name in ("mean_mlp", "last_mlp", "wa_mlp"):
            return code, rep_dim, 1, 1
        if name in ("mean_tr", "last_tr", "wa_tr"):
            return code, rep_dim, int(parts[-1]), 1
        if name == "wa_down_tr":
            return code, int(parts[-2]), int(parts[-1]), 1
        if name == "lw_tr":
            return code, rep_dim, int(parts[-2]), int(parts[-1])
        return code, int(parts[-3]), int(parts[-2]), int(parts[-1])
    except (ValueError, IndexError):
        raise ValueError(f"malformed TLTR mode {mode!r}") from None


class TLTR:
    def __init__(self, label_dim: int = 527, n_layer: int = 33, rep_dim: int = 1280, mode: str = "basic",
                 precision: str = "bf16", max_batch: int = 128):
        self.mode, self.n_layer, self.rep_dim, self.label_dim = mode, n_layer, rep_dim, label_dim
        code, inter, nt, nl = parse_mode(mode, rep_dim)
        self.inter_rep_dim = inter
        cfg = _lib.WatHeadConfig(rep_dim, n_layer, inter, label_dim, code, nt, nl,
                                 _lib.WAT_BF16 if precision == "bf16" else _lib.WAT_FP32, max_batch)
        h = C.c_void_p()
        self._device = torch.device("cuda", torch.cuda.current_device())      # the handle is bound to the device current at creation
        _lib.check(_lib.lib().wat_head_create(C.byref(cfg), C.byref(h)))
        self._h = h
        self._loaded = False

    def load_state_dict(self, sd: Dict[str, torch.Tensor], strict: bool = True) -> None:
        L = _lib.lib()
        for key, t in sd.items():
            a = t.detach().to(torch.float32).cpu().contiguous()
            _lib.check(L.wat_set_weight(self._h, key.encode(), C.c_void_p(a.data_ptr()), C.c_int64(a.numel())))
        _lib.check(L.wat_finalize(self._h))          # fails with "Missing key in state_dict: ..." like strict loading
        self._loaded = True

    @torch.no_grad()
    def forward(self, audio_rep: torch.Tensor) -> torch.Tensor:
        if not self._loaded:
            raise RuntimeError("TLTR: load_state_dict() first")
        assert audio_rep.ndim == 4 and audio_rep.shape[1] == self.n_layer and audio_rep.shape[3] == self.rep_dim, \
            "audio_rep must be [B, n_layer, T', rep_dim]"
        dev = audio_rep.device
        with torch.cuda.device(self._device):
            x = audio_rep.to(device=self._device, dtype=torch.float32).contiguous()
            B, _, Tp, _ = x.shape
            out = torch.empty(B, self.label_dim, device=self._device, dtype=torch.float32)
            _lib.check(_lib.lib().wat_head_forward(self._h, C.c_void_p(x.data_ptr()), B, Tp, C.c_void_p(out.data_ptr()),
                                                   C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        return out.to(dev)

    __call__ = forward

    def eval(self):
        return self

    def close(self) -> None:
        if getattr(self, "_h", None):
            _lib.lib().wat_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
