"""CPU: the oracle restatement (oracle/wat_oracle.py) against the fixtures produced by the REAL reference
(oracle/make_golden.py -> tests/golden/*.npz).  Runs anywhere; /root/reference is not needed."""
import os

import numpy as np
import pytest
import torch

import wat_oracle as O
from whisper_at import synth

MEL_COLS = np.r_[0:40, 1480:1520, 2960:3000, 40:2960:73]


def load(golden_dir, tag):
    z = np.load(os.path.join(golden_dir, tag + ".npz"), allow_pickle=True)
    n_mels, d, h, L, low, seed = [int(v) for v in z["meta"]]
    sd = synth.synth_state_dict(n_mels, d, L, bool(low), seed=seed, init=str(z["init"]))
    return z, sd, n_mels, h


def test_filterbank_properties():
    fb = O.mel_filterbank(80)
    assert fb.shape == (80, 201) and fb.dtype == np.float32
    assert (fb != 0).sum() == 391                       # measured on the reference asset
    assert not fb[:, 0].any() and not fb[:, 200].any()
    fb128 = O.mel_filterbank(128)
    assert not fb128[:, 0].any() and not fb128[:, 200].any()


def test_mel_known_answers(golden_dir):
    z = np.load(os.path.join(golden_dir, "mel_kat.npz"))
    assert torch.all(O.log_mel_clip(torch.zeros(480000)) == float(z["silence_value"]))
    t = torch.arange(480000) / 16000.0
    m = O.log_mel_clip(0.5 * torch.sin(2 * np.pi * 1000.0 * t))
    assert int(m[:, 100].argmax()) == int(z["tone1k_argmax_bin"])
    np.testing.assert_allclose(m[:, 100].numpy(), z["tone1k_col100"], atol=2e-6)
    assert float(m.max() - m.min()) <= 2.0 + 1e-6        # the reference's own test_audio.py:19 property
    short = synth.synth_clip(2)[:80000]
    np.testing.assert_allclose(O.log_mel_clip(short)[:, MEL_COLS].numpy(), z["short5s_cols"], atol=2e-6)
    # the reference's fp32 FFT is this far from the fp64 truth; the oracle's fp64 path reproduces that gap
    clip = synth.synth_clip(1)
    gap = (O.log_mel_clip(clip).double() - O.log_mel_clip(clip, dtype=torch.float64, explicit_dft=True)).abs().max()
    assert abs(float(gap) - float(z["ref_vs_fp64_maxabs"])) < 5e-6


@pytest.mark.parametrize("tag,clips,resolutions,starts", [
    ("tiny_default", (0, 1, 7), (10, 2, 0.4, 4, 30), (0,)),
    ("tiny_lively", (1, 7), (10, 4), (0, 10)),
    ("tiny_low_lively", (1,), (10, 2, 4), (0,)),
    ("base_lively", (0, 1), (10,), (0,)),
])
def test_oracle_matches_reference_fixtures(golden_dir, tag, clips, resolutions, starts):
    z, sd, n_mels, h = load(golden_dir, tag)
    for ci in clips:
        clip = synth.synth_clip(ci)
        mel = O.log_mel_clip(clip, n_mels)
        if f"mel_c{ci}" in z:
            np.testing.assert_allclose(mel[:, MEL_COLS].numpy(), z[f"mel_c{ci}"], atol=2e-6)
            assert abs(float(mel.max()) - float(z[f"mel_c{ci}_max"])) < 1e-6
        pooled = O.encoder_pooled(mel[None], sd, h)
        if f"pooled_c{ci}" in z:
            ref = z[f"pooled_c{ci}"]
            np.testing.assert_allclose(pooled[0][:, ::5, ::3].numpy(), ref, atol=2e-4 * max(1.0, np.abs(ref).max()))
        for res in resolutions:
            for a0 in starts:
                lg = O.tltr_head(pooled[:, :, a0:, :], sd, res)[0]
                ref = z[f"logits_c{ci}_r{res}_a{a0}"]
                assert lg.shape == ref.shape
                np.testing.assert_allclose(lg.numpy(), ref, atol=2e-4)


@pytest.mark.skipif(not os.environ.get("WAT_SLOW"), reason="set WAT_SLOW=1 for the small/medium/large fixtures (minutes of CPU)")
@pytest.mark.parametrize("tag,resolutions", [("small_low_lively", (2, 10)), ("medium_low_lively", (10,)),
                                             ("large_v2_m128_lively", (10,))])
def test_oracle_matches_reference_fixtures_large(golden_dir, tag, resolutions):
    z, sd, n_mels, h = load(golden_dir, tag)
    clip = synth.synth_clip(1)
    for res in resolutions:
        lg = O.tag(clip[None], sd, h, n_mels, res)[0]
        np.testing.assert_allclose(lg.numpy(), z[f"logits_c1_r{res}_a0"], atol=2e-4)


def test_batch_semantics_are_per_clip():
    """oracle batch of 2 == two single-clip runs (the reference itself drops clips 1.. of a batch)."""
    sd = synth.synth_state_dict(80, 384, 4, False, seed=1, init="lively")
    a = synth.synth_batch(2, start=1)
    both = O.tag(a, sd, 6, 80, 10)
    for i in range(2):
        one = O.tag(a[i:i + 1], sd, 6, 80, 10)
        assert float((both[i] - one[0]).abs().max()) < 1e-5


def test_decision_window_arithmetic():
    assert [O.decision_window(r) for r in (10, 2, 0.4, 4, 30, 0.8)] == [25, 5, 1, 10, 75, 2]


def _tltr_cases(golden_dir):
    z = np.load(os.path.join(golden_dir, "tltr_variants.npz"))
    for i, c in enumerate(z["cases"]):
        mode, L, T, d, B, nc = str(c).split("|")
        yield i, mode, int(L), int(T), int(d), int(B), int(nc), z[f"case{i}"]


def test_oracle_tltr_variants_match_reference_fixtures(golden_dir):
    """every head of the training recipe's TLTR class (src/whisper_at_train/models.py:108-200; fixtures made from the
    real class by oracle/make_golden_tltr.py) — the large-v2-sized case is left to WAT_SLOW"""
    from whisper_at import synth
    n = 0
    for i, mode, L, T, d, B, nc, ref in _tltr_cases(golden_dir):
        if d > 512 and not os.environ.get("WAT_SLOW"):
            continue
        sd = synth.synth_tltr_state_dict(mode, L, d, nc, seed=1)
        x = synth.synth_audio_rep(B, L, T, d, seed=7 + i)
        with torch.no_grad():
            got = O.tltr_variant(x, sd, mode)
        assert got.shape == ref.shape
        assert float((got - torch.from_numpy(ref)).abs().max()) < 2e-5, mode
        n += 1
    assert n >= 10
