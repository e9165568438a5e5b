"""Golden vectors for the TL-TR head variants of the training recipe, from the REAL reference class.

Run in the build container only (needs /root/reference, read-only):

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_tltr.py

For every mode string it builds the reference `TLTR(label_dim, n_layer, rep_dim, mode)`
(/root/reference/src/whisper_at_train/models.py), loads the seeded weights of whisper_at/synth.py with strict=True (so
synth's key names and shapes are pinned to the reference's), runs the reference forward on the seeded [B, L, T', d] input,
asserts oracle/wat_oracle.py:tltr_variant reproduces it, and stores the reference logits in tests/golden/tltr_variants.npz.
"""
from __future__ import annotations

import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_PKG = "/root/reference/package/whisper-at"
REF_TRAIN = "/root/reference/src/whisper_at_train/models.py"
sys.dont_write_bytecode = True
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, REF_PKG)

import wat_oracle as O  # noqa: E402

_spec = importlib.util.spec_from_file_location("wat_synth", os.path.join(ROOT, "whisper-at_b200", "whisper_at", "synth.py"))
synth = importlib.util.module_from_spec(_spec)
sys.modules["wat_synth"] = synth
_spec.loader.exec_module(synth)

# the training code imports `whisper.model` (openai-whisper); the reference package carries the same classes
import whisper_at as R  # noqa: E402
import whisper_at.model as R_model  # noqa: E402
sys.modules.setdefault("whisper", R)
sys.modules.setdefault("whisper.model", R_model)
_mspec = importlib.util.spec_from_file_location("ref_train_models", REF_TRAIN)
ref_models = importlib.util.module_from_spec(_mspec)
_mspec.loader.exec_module(ref_models)

# (mode, n_layer, T', rep_dim, batch[, label_dim])   label_dim 533 = the SONYC fine-tune's extended classifier (run.py:148-184)
CASES = [
    ("mean_mlp", 4, 25, 384, 3), ("last_mlp", 4, 25, 384, 3), ("wa_mlp", 4, 25, 384, 3),
    ("mean_tr_1", 4, 25, 384, 3), ("last_tr_4", 4, 25, 384, 3), ("wa_tr_1", 6, 25, 512, 2),
    ("wa_down_tr_256_1", 4, 25, 384, 3), ("lw_tr_1_8", 4, 25, 384, 2), ("lw_down_tr_256_1_8", 6, 25, 512, 2),
    ("lw_tr_1_8", 32, 25, 1280, 1), ("mean_tr_1", 4, 10, 384, 2), ("lw_down_tr_256_1_8", 4, 25, 384, 2, 533),
]


def main():
    torch.set_num_threads(os.cpu_count())
    out = {}
    for i, case in enumerate(CASES):
        mode, L, T, d, B = case[:5]
        nc = case[5] if len(case) > 5 else 527
        sd = synth.synth_tltr_state_dict(mode, L, d, nc, seed=1)
        m = ref_models.TLTR(label_dim=nc, n_layer=L, rep_dim=d, mode=mode).eval()
        m.load_state_dict(sd, strict=True)
        x = synth.synth_audio_rep(B, L, T, d, seed=7 + i)
        with torch.no_grad():
            ref = m(x).float()
            mine = O.tltr_variant(x, sd, mode)
        err = (ref - mine).abs().max().item()
        assert ref.shape == (B, nc) and err < 2e-5, (mode, ref.shape, err)
        print(f"{mode:22s} L={L:2d} T'={T} d={d:4d} B={B}: oracle vs reference max|d| = {err:.2e}, logit std {ref.std():.3f}")
        out[f"case{i}"] = ref.numpy()
    out["cases"] = np.array(["|".join(str(v) for v in (list(c) + [527])[:6]) for c in CASES])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "tltr_variants.npz"), **out)
    print("wrote tests/golden/tltr_variants.npz")


if __name__ == "__main__":
    main()
