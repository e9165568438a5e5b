"""whisper_at (B200): drop-in for the audio-tagging path of chat-prompt/whisper-at.

Public names follow package/whisper-at/whisper_at/__init__.py: load_model, available_models, log_mel_spectrogram,
pad_or_trim, load_audio, parse_at_label, print_label_name, print_support_language, ModelDimensions, Whisper,
transcribe.  Compute is done by libwat.so (hand-written sm_100a CUDA, see ../csrc and include/wat.h)."""
from __future__ import annotations

import hashlib
import io
import os
import urllib.parse
import urllib.request
import warnings
from typing import List, Optional, Union

import torch

from .at_post_processing import parse_at_label, print_label_name, print_support_language
from .audio import load_audio, log_mel_spectrogram, pad_or_trim
from .model import ModelDimensions, Whisper
from .transcribe import transcribe
from .version import __version__

# the public checkpoints of the reference (__init__.py:18-51): OpenAI Whisper weights + the TL-TR heads
# Where the public checkpoints live (the reference's __init__.py:18-51 lists the same files): OpenAI Whisper weights plus
# the TL-TR heads trained by the Whisper-AT authors.  The registry is data, kept in assets/checkpoints.json.
def _registry():
    import json
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets", "checkpoints.json")) as f:
        reg = json.load(f)
    whisper_urls = {name: reg["openai_prefix"] + tail for name, tail in reg["openai"].items()}
    head_urls = {name: reg["at_prefix"] + ident + reg["at_suffix"] for name, ident in reg["at"].items()}
    return whisper_urls, head_urls


_MODELS, _MODELS_AT = _registry()


def available_models() -> List[str]:
    """Names accepted by load_model (__init__.py:115)."""
    return list(_MODELS.keys())


def _fetch(url: str, root: str, in_memory: bool) -> Union[bytes, str]:
    """Cached download with the reference's cache-file naming (__init__.py:68-112): basename of the URL *path*
    ("tiny_ori.pth" for ".../tiny_ori.pth?dl=1"), so caches written by the reference are reused.  The OpenAI URLs carry
    the file's sha256 as their second-to-last path component; it is checked (the reference has the check commented out)."""
    os.makedirs(root, exist_ok=True)
    path = urllib.parse.urlparse(url).path
    target = os.path.join(root, os.path.basename(path))
    parts = path.split("/")
    expected_sha = parts[-2] if len(parts) >= 2 and len(parts[-2]) == 64 and all(c in "0123456789abcdef" for c in parts[-2]) else None
    if os.path.exists(target) and not os.path.isfile(target):
        raise RuntimeError(f"{target} exists and is not a regular file")
    fresh = not os.path.isfile(target)
    if fresh:
        try:
            with urllib.request.urlopen(url) as src, open(target + ".part", "wb") as dst:
                while True:
                    buf = src.read(1 << 20)
                    if not buf:
                        break
                    dst.write(buf)
            os.replace(target + ".part", target)
        except Exception as e:
            raise RuntimeError(f"could not download {url} to {target} ({e}); place the file there manually") from e
    if expected_sha is not None:
        hsh = hashlib.sha256()
        with open(target, "rb") as f:
            for buf in iter(lambda: f.read(1 << 20), b""):
                hsh.update(buf)
        if hsh.hexdigest() != expected_sha:
            if fresh:
                os.remove(target)
                raise RuntimeError(f"{url} has been downloaded but its SHA256 checksum does not match; retry loading the model")
            # a file that was already in the cache is used as it is, like the reference does (its check is commented out)
            warnings.warn(f"{target} exists, but its SHA256 checksum does not match the one in {url}")
    if in_memory:
        with open(target, "rb") as f:
            return f.read()
    return target


def _torch_load(src, device):
    with (io.BytesIO(src) if isinstance(src, bytes) else open(src, "rb")) as fp:
        return torch.load(fp, map_location=device, weights_only=True)      # checkpoints are plain dicts of tensors / ints


def load_model(name: str, device: Optional[Union[str, torch.device]] = None, download_root: str = None,
               in_memory: bool = False, at_low_compute=False, *, at_checkpoint: Optional[str] = None,
               precision: str = "bf16") -> Whisper:
    """Load a Whisper-AT model (__init__.py:120-196).

    name: an official model name, or the path of an OpenAI-format checkpoint {"dims", "model_state_dict"}; in the
    second case pass the TL-TR head weights with `at_checkpoint` (the reference cannot load a head for a path).
    at_low_compute selects the TL-TR-512 head (`<name>_low`), as in the reference.  `precision` ("bf16" | "fp32")
    is this implementation's compute mode."""
    if device is None:
        device = "cuda" if torch.cuda.is_available() else "cpu"
    if download_root is None:
        default = os.path.join(os.path.expanduser("~"), ".cache")
        download_root = os.path.join(os.getenv("XDG_CACHE_HOME", default), "whisper")
    at_name = name + "_low" if at_low_compute else name
    if name in _MODELS:
        if at_name not in _MODELS_AT:
            raise KeyError(at_name)                              # same failure as the reference's dict lookup
        ckpt_src = _fetch(_MODELS[name], download_root, in_memory)
        at_src = _fetch(_MODELS_AT[at_name], download_root, in_memory)
    elif os.path.isfile(name):
        ckpt_src = open(name, "rb").read() if in_memory else name
        if at_checkpoint is None or not os.path.isfile(at_checkpoint):
            raise RuntimeError("loading from a checkpoint path needs at_checkpoint=<path of the TL-TR .pth>")
        at_src = open(at_checkpoint, "rb").read() if in_memory else at_checkpoint
    else:
        raise RuntimeError(f"Model {name} not found; available models = {available_models()}")
    checkpoint = _torch_load(ckpt_src, "cpu")
    checkpoint_at = _torch_load(at_src, "cpu")
    dims = ModelDimensions(**checkpoint["dims"])
    model = Whisper(dims, at_low_compute=at_low_compute, precision=precision)
    combined = {}
    combined.update(checkpoint["model_state_dict"])
    combined.update(checkpoint_at)
    model.load_state_dict(combined, strict=True)
    return model.to(device)


Whisper.transcribe = transcribe

__all__ = ["load_model", "available_models", "log_mel_spectrogram", "pad_or_trim", "load_audio", "parse_at_label",
           "print_label_name", "print_support_language", "ModelDimensions", "Whisper", "transcribe", "__version__"]
