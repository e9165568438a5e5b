"""bench.py infrastructure checked on the CPU: the PyTorch-eager GPU baseline (baseline/torch_eager.py) is a faithful
implementation of the path (against the oracle, fp32, both attention variants, both head kinds), and the bench presets are
BASELINE.json's configurations."""
import importlib.util
import os
import sys

import pytest
import torch

import wat_oracle as O
from whisper_at import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "baseline"))


@pytest.mark.parametrize("low", [False, True])
@pytest.mark.parametrize("attention", ["sdpa", "materialized"])
def test_torch_eager_baseline_matches_the_oracle(low, attention):
    from torch_eager import EagerTagger
    d, h, L = synth.MODEL_SHAPES["tiny"]
    sd = synth.synth_state_dict(80, d, L, low, seed=1, init="lively")
    audio = synth.synth_batch(2, start=1)
    tagger = EagerTagger(sd, h, 80, "cpu", torch.float32, attention)
    for res in (10, 4):
        ref = O.tag(audio, sd, h, 80, res)
        got = tagger.tag(audio, res)
        assert got.shape == ref.shape
        assert float((got - ref).abs().max()) <= 1e-4
    # chunked calls give the same logits (bench.py runs the materialised-qk variant 32 clips at a time)
    assert torch.allclose(tagger.tag(audio, 10, chunk=1), tagger.tag(audio, 10), atol=1e-5)


def test_bench_presets_are_the_baseline_configs():
    spec = importlib.util.spec_from_file_location("wat_bench", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    for i, cfg in enumerate(synth.BASELINE_CONFIGS, start=1):
        name, n_mels, low, res, total, per_gpu = bench.CONFIGS[i]
        assert (name, n_mels, low, res, total) == (cfg.name, cfg.n_mels, cfg.low, float(cfg.at_time_res), cfg.batch)
        assert per_gpu <= total
    fl = bench.flops_per_clip(1280, 32, 128, False, 10)
    assert abs(fl["encoder"] / 2273.8e9 - 1) < 2e-3                 # large-v2 (128 mel) encoder total
    assert abs((fl["total"] - fl["encoder"]) / 98.5e9 - 1) < 1e-2   # head @ res 10, full TL-TR


def test_cpu_arm_uses_the_reference_itself_when_present():
    """bench.py's CPU legs run the unmodified reference package from oracle/_ref (placed by oracle/make_ref.py, which
    needs /root/reference) and check it against the oracle port on the first clip; without oracle/_ref they fall back
    to the port and label the line "port"."""
    spec = importlib.util.spec_from_file_location("wat_bench", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    have_ref = os.path.isfile(os.path.join(ROOT, "oracle", "_ref", "whisper_at", "model.py"))
    arm = bench.cpu_arm("tiny", 80, False, 10.0, 0.0, 1, 4)
    assert len(arm["times"]) == 1 and arm["times"][0] > 0
    if have_ref:
        assert arm["kind"] == "reference" and arm["port_max_abs_diff"] <= 2e-4
        # low-compute head and another resolution, straight through the runner
        times, lg = bench.reference_time_clips("tiny", 80, True, 4.0, 1, 4)
        d, h, L = synth.MODEL_SHAPES["tiny"]
        sd = synth.synth_state_dict(80, d, L, True, seed=1, init="lively")
        ref = O.tag(synth.synth_clip(1)[None], sd, h, 80, 4.0)[0]
        assert lg.shape == ref.shape and float((lg - ref).abs().max()) <= 2e-4
    else:
        assert arm["kind"] == "port"
