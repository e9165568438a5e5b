"""CPU: host-side logic of the drop-in package and the C-ABI surface (no compute calls)."""
import ctypes as C
import os
import re
import subprocess
import warnings

import numpy as np
import pytest
import torch

import whisper_at
from whisper_at import _lib, synth
from whisper_at.transcribe import check_at_time_res

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "wat.h")).read()
    declared = set(re.findall(r"WAT_API\s+[\w\s\*]+?\b(wat_\w+)\s*\(", hdr))
    assert len(declared) >= 15
    L = _lib.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert declared == set(_lib.EXPORTS)
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (wat_\w+)", out))
    assert declared <= exported
    assert L.wat_abi_version() == 2


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_no_cpu_fallback():
    L = _lib.lib()
    cfg = _lib.WatConfig(80, 1500, 384, 6, 4, 0, 0, 527, 0, 0)
    h = C.c_void_p()
    assert L.wat_create(C.byref(cfg), C.byref(h)) == _lib.WAT_ERR_CUDA
    assert b"no CPU fallback" in L.wat_last_error()
    dims = whisper_at.ModelDimensions(80, 1500, 384, 6, 4, 51865, 448, 384, 6, 4)
    m = whisper_at.Whisper(dims)
    with pytest.raises(RuntimeError, match="CUDA only"):
        m.encoder(torch.zeros(1, 80, 3000))
    with pytest.raises(RuntimeError, match="CUDA"):
        whisper_at.log_mel_spectrogram(np.zeros(16000, np.float32))


def test_config_validation_messages():
    L = _lib.lib()
    h = C.c_void_p()
    bad = _lib.WatConfig(64, 1500, 384, 6, 4, 0, 0, 527, 0, 0)
    assert L.wat_create(C.byref(bad), C.byref(h)) == _lib.WAT_ERR_INVALID
    assert b"Unsupported n_mels: 64" in L.wat_last_error()          # the reference's assert text (audio.py:103)
    bad = _lib.WatConfig(80, 1500, 400, 6, 4, 0, 0, 527, 0, 0)
    assert L.wat_create(C.byref(bad), C.byref(h)) == _lib.WAT_ERR_INVALID


def test_state_dict_keys_match_reference_names():
    for low in (False, True):
        dims = whisper_at.ModelDimensions(80, 1500, 384, 6, 4, 51865, 448, 384, 6, 4)
        m = whisper_at.Whisper(dims, at_low_compute=low)
        sd = synth.synth_state_dict(80, 384, 4, low)
        assert set(m.state_dict()) == set(sd) | {"encoder.positional_embedding"}
        for k, v in sd.items():
            assert tuple(m.state_dict()[k].shape) == tuple(v.shape), k
        # decoder.* keys of a full OpenAI checkpoint are accepted and dropped; strict otherwise
        full = dict(sd)
        full["decoder.token_embedding.weight"] = torch.zeros(4, 4)
        full["encoder.positional_embedding"] = synth.sinusoid_table(1500, 384)
        m.load_state_dict(full, strict=True)
        bad = dict(sd)
        bad.pop("at_model.mlp_layer.1.bias")
        with pytest.raises(RuntimeError):
            m.load_state_dict(bad, strict=True)
    assert m.is_multilingual and m.device == torch.device("cpu")
    # the TL-TR parameter counts the reference publishes (README.md:258-269)
    def at_params(name, low):
        d, _, L = synth.MODEL_SHAPES[name]
        return sum(int(np.prod(s)) for k, s in synth.tagging_state_shapes(80, d, L, low).items() if k.startswith("at_model."))
    assert abs(at_params("large-v2", False) / 1e6 - 40.03) < 0.01
    assert abs(at_params("small", False) / 1e6 - 14.581) < 0.001
    assert abs(at_params("small", True) / 1e6 - 6.970) < 0.001


def test_load_model_errors():
    with pytest.raises(RuntimeError, match="Model nonexistent not found; available models"):
        whisper_at.load_model("nonexistent")
    assert whisper_at.available_models()[0] == "tiny.en" and "large-v2" in whisper_at.available_models()
    with pytest.raises(KeyError):
        whisper_at.load_model("tiny", at_low_compute=True, download_root="/tmp/wat_none")   # no tiny_low head exists


def test_at_time_res_validation_matches_reference():
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        assert check_at_time_res(10) == 1000                      # no warning at the trained resolution
    for bad in (0.5, 1, 4.4, 8.8):                                # 4.4*100 == 440.00000000000006
        with pytest.raises(AssertionError, match="must be an integer multiple of 0.4 second"):
            check_at_time_res(bad)
    with pytest.warns(UserWarning, match="Current at_time_res is 2.00 second"):
        assert check_at_time_res(2) == 200
    with pytest.warns(UserWarning):
        check_at_time_res(0.4)


def test_pad_or_trim():
    a = np.arange(10, dtype=np.float32)
    assert whisper_at.pad_or_trim(a, 4).tolist() == [0, 1, 2, 3]
    assert whisper_at.pad_or_trim(a, 12).tolist() == list(range(10)) + [0, 0]
    t = torch.arange(12.).reshape(2, 6)
    assert whisper_at.pad_or_trim(t, 4).shape == (2, 4)
    p = whisper_at.pad_or_trim(t, 8)
    assert p.shape == (2, 8) and p[:, 6:].abs().sum() == 0 and torch.equal(p[:, :6], t)
    p0 = whisper_at.pad_or_trim(t, 3, axis=0)
    assert p0.shape == (3, 6) and torch.equal(p0[:2], t)
    assert whisper_at.pad_or_trim(np.zeros((80, 2000)), 3000).shape == (80, 3000)


def test_parse_at_label_matches_reference_fixture(golden_dir):
    z = np.load(os.path.join(golden_dir, "api_tiny.npz"), allow_pickle=True)
    g = torch.Generator().manual_seed(5)
    fixed = torch.randn(4, 527, generator=g)
    res = {"language": "en", "at_time_res": 4.8, "audio_tag": fixed}
    p = whisper_at.parse_at_label(res, language="follow_asr", top_k=3, p_threshold=1.5, include_class_list=list(range(0, 527, 2)))
    assert repr(p) == str(z["parse_fixed"])
    p2 = whisper_at.parse_at_label(res, language="zh", top_k=2, p_threshold=-1)
    assert repr(p2) == str(z["parse_fixed_zh"])
    with pytest.warns(UserWarning, match="language not supported"):
        p3 = whisper_at.parse_at_label(dict(res, language="xx"), top_k=1)
    assert p3[0]["audio tags"][0][0] == whisper_at.parse_at_label(res, language="en", top_k=1)[0]["audio tags"][0][0]


def test_label_assets(capsys):
    whisper_at.print_label_name("en")
    out = capsys.readouterr().out.splitlines()
    assert len(out) == 527 and out[0] == "index: 0 : Speech"
    whisper_at.print_support_language()
    out = capsys.readouterr().out.splitlines()
    assert len(out) == 85 and out[0] == "language code: en : english"


def test_synth_is_deterministic():
    a, b = synth.synth_clip(3), synth.synth_clip(3)
    assert torch.equal(a, b) and a.shape == (480000,) and a.dtype == torch.float32
    assert synth.synth_clip(7)[-160000:].abs().max() == 0
    s1 = synth.synth_state_dict(80, 384, 4, False, seed=1, init="lively")
    s2 = synth.synth_state_dict(80, 384, 4, False, seed=1, init="lively")
    assert all(torch.equal(s1[k], s2[k]) for k in s1)
    assert not torch.equal(s1["encoder.conv1.weight"], synth.synth_state_dict(80, 384, 4, False, seed=2, init="lively")["encoder.conv1.weight"])


def _fake_checkpoints(tmp_path, low):
    """An OpenAI-format checkpoint {"dims", "model_state_dict"} (with decoder.* tensors, as the real files have) and the
    flat AT .pth whose keys already start with at_model. (__init__.py:172-191)."""
    dims = dict(n_mels=80, n_audio_ctx=1500, n_audio_state=384, n_audio_head=6, n_audio_layer=4, n_vocab=51865,
                n_text_ctx=448, n_text_state=384, n_text_head=6, n_text_layer=4)
    sd = synth.synth_state_dict(80, 384, 4, low, seed=3, init="lively")
    enc = {k: v for k, v in sd.items() if k.startswith("encoder.")}
    enc["encoder.positional_embedding"] = synth.sinusoid_table(1500, 384)
    enc["decoder.token_embedding.weight"] = torch.zeros(8, 384)
    enc["decoder.positional_embedding"] = torch.zeros(448, 384)
    at = {k: v for k, v in sd.items() if k.startswith("at_model.")}
    return dims, sd, enc, at


@pytest.mark.parametrize("low", [False, True])
def test_load_model_from_checkpoint_paths(tmp_path, low):
    dims, sd, enc, at = _fake_checkpoints(tmp_path, low)
    ck, atp = tmp_path / "enc.pt", tmp_path / "at.pth"
    torch.save({"dims": dims, "model_state_dict": enc}, ck)
    torch.save(at, atp)
    m = whisper_at.load_model(str(ck), device="cpu", at_low_compute=low, at_checkpoint=str(atp))
    assert isinstance(m, whisper_at.Whisper) and m.at_low_compute == low
    assert m.at_model.mode == ("tl_down_tr_512_1_8" if low else "tl_tr_1_8")
    got = m.state_dict()
    for k, v in sd.items():
        assert torch.equal(got[k], v), k
    m2 = whisper_at.load_model(str(ck), device="cpu", at_low_compute=low, at_checkpoint=str(atp), in_memory=True)
    assert all(torch.equal(m2.state_dict()[k], v) for k, v in sd.items())
    with pytest.raises(RuntimeError, match="at_checkpoint"):
        whisper_at.load_model(str(ck), device="cpu")
    # a head that does not match at_low_compute is rejected by the strict load, as in the reference
    with pytest.raises(RuntimeError):
        whisper_at.load_model(str(ck), device="cpu", at_low_compute=not low, at_checkpoint=str(atp))


def test_load_model_by_name_uses_the_reference_cache_layout(tmp_path):
    """Official names resolve to <download_root>/<basename(urlparse(url).path)> exactly as the reference's _download caches
    them (__init__.py:70-75: "tiny.pt", "tiny_ori.pth" without the "?dl=1" query), so a cache populated by the reference
    is picked up without network access.  A cached file whose SHA256 differs from the one in the OpenAI URL is still
    used (the reference's check is commented out) but warned about."""
    dims, sd, enc, at = _fake_checkpoints(tmp_path, False)
    torch.save({"dims": dims, "model_state_dict": enc}, tmp_path / "tiny.pt")
    torch.save(at, tmp_path / "tiny_ori.pth")
    with pytest.warns(UserWarning, match="SHA256 checksum does not match"):
        m = whisper_at.load_model("tiny", device="cpu", download_root=str(tmp_path))
    assert all(torch.equal(m.state_dict()[k], v) for k, v in sd.items())
    assert m.dims.n_audio_layer == 4 and m.precision == "bf16"
    with pytest.raises(RuntimeError, match="could not download"):
        whisper_at.load_model("base", device="cpu", download_root=str(tmp_path / "empty"))


def test_feature_file_format_matches_reference_loader(tmp_path):
    """the npz written for the TL-TR training recipe is what dataloader_feat.py:97-125 reads: key arr_0, [L, T', d],
    padded / cut to 25 steps by the loader"""
    from whisper_at import features
    feat = np.random.default_rng(0).standard_normal((4, 25, 384)).astype(np.float32)
    path = str(tmp_path / "clip.npz")
    features.save_feature_npz(path, feat)
    z = np.load(path)
    assert z.files == ["arr_0"] and z["arr_0"].dtype == np.float32
    assert np.array_equal(features.load_feature_npz(path), feat)
    t = torch.Tensor(z["arr_0"])                                   # the reference loader's next steps
    t = t[:, :25, :] if t.shape[1] >= 25 else torch.nn.functional.pad(t, (0, 0, 0, 25 - t.shape[1]))
    assert tuple(t.shape) == (4, 25, 384)


def test_tltr_mode_strings():
    """mode-string grammar of the training recipe's TLTR class (models.py:56-106)"""
    from whisper_at.tltr import parse_mode
    from whisper_at import _lib as L, synth
    assert parse_mode("mean_mlp", 1280) == (L.HEAD_MODES["mean_mlp"], 1280, 1, 1)
    assert parse_mode("last_tr_4", 768) == (L.HEAD_MODES["last_tr"], 768, 4, 1)
    assert parse_mode("wa_down_tr_512_1", 1280) == (L.HEAD_MODES["wa_down_tr"], 512, 1, 1)
    assert parse_mode("lw_tr_1_8", 1280) == (L.HEAD_MODES["lw_tr"], 1280, 1, 8)
    assert parse_mode("lw_down_tr_512_1_8", 1280) == (L.HEAD_MODES["lw_down_tr"], 512, 1, 8)
    with pytest.raises(ValueError):
        parse_mode("basic", 1280)
    with pytest.raises(ValueError):
        parse_mode("mean_tr_x", 1280)
    # synth's key set for a mode is what the reference module exposes (pinned by strict loading in make_golden_tltr.py)
    keys = set(synth.tltr_state_shapes("wa_down_tr_256_1", 4, 384))
    assert {"layer_weight", "down_layer.1.weight", "time_tr.attn.key.weight", "mlp_layer.1.bias"} <= keys
    assert not any(k.startswith("layer_tr") for k in keys)


def test_bench_flop_model_matches_survey_table():
    """bench.py's algorithmic FLOPs per clip (the numerator of roofline.achieved) against SURVEY.md §8d's table"""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("wat_bench", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    for (d, L, n_mels, low, res), enc_g, head_g in [((384, 4, 80, False, 10), 36.9, 1.12), ((512, 6, 80, False, 10), 87.4, 2.97),
                                                     ((768, 12, 80, True, 10), 344.2, 6.65), ((1024, 24, 80, True, 10), 1138.1, 13.8),
                                                     ((1280, 32, 128, False, 10), 2273.8, 98.5), ((1280, 32, 80, False, 0.4), 2272.7, 189.0)]:
        f = bench.flops_per_clip(d, L, n_mels, low, res)
        assert abs(f["encoder"] / 1e9 - enc_g) < 0.002 * enc_g + 0.06, (d, f["encoder"] / 1e9)
        assert abs((f["total"] - f["encoder"]) / 1e9 - head_g) < 0.01 * head_g + 0.02, (d, (f["total"] - f["encoder"]) / 1e9)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs next to ours): one JSON line with the contract's keys"""
    import json
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--model", "tiny", "--steps", "1",
                        "--warmup", "1"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    j = json.loads(p.stdout.strip().splitlines()[-1])
    assert j["impl"] == "reference" and j["unit"] == "audio-s/s" and j["higher_is_better"] is True and j["value"] > 0
    assert j["e2e"] == {"value": j["value"], "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    have_ref = os.path.isfile(os.path.join(root, "oracle", "_ref", "whisper_at", "model.py"))     # placed by oracle/make_ref.py
    assert j["cpu_baseline"]["kind"] == ("reference" if have_ref else "port")
    assert j["cpu_baseline"]["cores"] >= 1 and j["cpu_baseline"]["value"] == j["value"]
    if have_ref:
        assert j["cpu_baseline"]["port_max_abs_diff"] <= 2e-4
    assert "workload" in j["config"] and j["vs_baseline"] is None
