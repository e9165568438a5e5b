"""whisper_at (B200): drop-in for the audio-tagging path of chat-prompt/whisper-at.

Public names follow package/whisper-at/whisper_at/__init__.py: load_model, available_models, log_mel_spectrogram,
pad_or_trim, load_audio, parse_at_label, print_label_name, print_support_language, ModelDimensions, Whisper,
transcribe.  Compute is done by libwat.so (hand-written sm_100a CUDA, see ../csrc and include/wat.h)."""
from __future__ import annotations

import io
import os
import urllib.request
from typing import List, Optional, Union

import torch

from .at_post_processing import parse_at_label, print_label_name, print_support_language
from .audio import load_audio, log_mel_spectrogram, pad_or_trim
from .model import ModelDimensions, Whisper
from .transcribe import transcribe
from .version import __version__

# the public checkpoints of the reference (__init__.py:18-51): OpenAI Whisper weights + the TL-TR heads
_OPENAI = "https://openaipublic.azureedge.net/main/whisper/models/"
_MODELS = {
    "tiny.en": _OPENAI + "d3dd57d32accea0b295c96e26691aa14d8822fac7d9d27d5dc00b4ca2826dd03/tiny.en.pt",
    "tiny": _OPENAI + "65147644a518d12f04e32d6f3b26facc3f8dd46e5390956a9424a650c0ce22b9/tiny.pt",
    "base.en": _OPENAI + "25a8566e1d0c1e2231d1c762132cd20e0f96a85d16145c3a00adf5d1ac670ead/base.en.pt",
    "base": _OPENAI + "ed3a0b6b1c0edf879ad9b11b1af5a0e6ab5db9205f891f668f8b0e6c6326e34e/base.pt",
    "small.en": _OPENAI + "f953ad0fd29cacd07d5a9eda5624af0f6bcf2258be67c92b79389873d91e0872/small.en.pt",
    "small": _OPENAI + "9ecf779972d90ba49c06d968637d720dd632c55bbf19d441fb42bf17a411e794/small.pt",
    "medium.en": _OPENAI + "d7440d1dc186f76616474e0ff0b3b6b879abc9d1a4926b7adfa41db2d497ab4f/medium.en.pt",
    "medium": _OPENAI + "345ae4da62f9b3d59415adc60127b97c714f32e89e936602e85993674d08dcb1/medium.pt",
    "large-v1": _OPENAI + "e4b87e7e0bf463eb8e6956e646f1e277e901512310def2c24bf0e11bd3c28e9a/large-v1.pt",
    "large-v2": _OPENAI + "81f7c96c852ee8fc832187b0132e569d6c3065a3252ed18e56effd0b6a73e524/large-v2.pt",
    "large": _OPENAI + "81f7c96c852ee8fc832187b0132e569d6c3065a3252ed18e56effd0b6a73e524/large-v2.pt",
}
_DROPBOX = "https://www.dropbox.com/s/"
_AT_IDS = {
    "tiny.en": "atq9so6w0qug5ai/tiny.en_ori", "tiny": "cib4q4iz6g758l0/tiny_ori",
    "base.en": "qtzgsbuquoz0afn/base.en_ori", "base": "2odwh42u6e9ger7/base_ori",
    "small.en": "cyx50ycl1ul7lji/small.en_ori", "small.en_low": "507o66zgl8v6ddd/small.en_low",
    "small": "jftj9s0kr4ycvr1/small_ori", "small_low": "a1x0416v58f7wrf/small_low",
    "medium.en": "bbvylvmgns8ja4p/medium.en_ori", "medium.en_low": "2q5wprr8f9gti5t/medium.en_low",
    "medium": "65aabayr7o819az/medium_ori", "medium_low": "0mnfmcasram4n6o/medium_low",
    "large-v1": "b8x2en1fdzc8nhk/large-v1_ori", "large-v1_low": "5o79h70wyla8jlk/large-v1_low",
    "large-v2": "3zxpyvdrxy22eq7/large-v2_ori", "large-v2_low": "jw2rh4uylhqgn85/large-v2_low",
    "large": "3zxpyvdrxy22eq7/large-v2_ori", "large_low": "jw2rh4uylhqgn85/large-v2_low",
}
_MODELS_AT = {k: f"{_DROPBOX}{v}.pth?dl=1" for k, v in _AT_IDS.items()}


def available_models() -> List[str]:
    """Names accepted by load_model (__init__.py:115)."""
    return list(_MODELS.keys())


def _fetch(url: str, root: str, in_memory: bool) -> Union[bytes, str]:
    """Cached download with the reference's cache-file naming (__init__.py:68-112): basename of the URL."""
    os.makedirs(root, exist_ok=True)
    target = os.path.join(root, os.path.basename(url))
    if os.path.exists(target) and not os.path.isfile(target):
        raise RuntimeError(f"{target} exists and is not a regular file")
    if not os.path.isfile(target):
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
    if in_memory:
        with open(target, "rb") as f:
            return f.read()
    return target


def _torch_load(src, device):
    with (io.BytesIO(src) if isinstance(src, bytes) else open(src, "rb")) as fp:
        return torch.load(fp, map_location=device)


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
