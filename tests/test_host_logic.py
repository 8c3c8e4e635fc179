"""CPU: host-side mirror of the reference interface -- module tree / state-dict names, config validation, weight-decay
split (ref:training/run_distillation.py:760-778,1386-1407), flat buffers, collator-shaped helpers."""
import numpy as np
import pytest
import torch
from torch import nn

from distil_whisper_b200.modeling import DistilWhisperB200ForConditionalGeneration, WhisperB200Config, sinusoids
from distil_whisper_b200.optim import FlatBuffers, decay_split, get_parameter_names
from distil_whisper_b200 import engine
from distil_whisper_b200.feature_extraction import WhisperFeatureExtractorB200, slaney_mel_filter_bank
from oracle import whisper_oracle as wo
from oracle import logmel_oracle as lo


def _tiny():
    return DistilWhisperB200ForConditionalGeneration(wo.PRESETS["tiny-student"].to_dict())


def test_state_dict_names_and_shapes_are_hf():
    m = _tiny()
    expected = {k: tuple(v) for k, v in wo.param_shapes(wo.PRESETS["tiny-student"]).items()}
    expected["proj_out.weight"] = expected["model.decoder.embed_tokens.weight"]
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == expected
    assert not m.model.encoder.embed_positions.weight.requires_grad
    assert torch.allclose(m.model.encoder.embed_positions.weight, wo.sinusoids(50, 128))
    assert float(m.model.decoder.embed_tokens.weight[500].abs().sum()) == 0.0          # padding_idx row
    sd = wo.init_state_dict(wo.PRESETS["tiny-student"], 3)
    m.load_hf_state_dict(sd)
    assert torch.equal(m.model.decoder.layers[0].fc1.weight, sd["model.decoder.layers.0.fc1.weight"])


def test_param_counts_match_readme():
    # README: 756 M (distil-large-v3), 1550 M (large-v3), 166 M (distil-small.en)  -- ref:README.md:13,15,18
    for name, expect in (("distil-large-v3", 756.4e6), ("large-v3", 1543.5e6), ("distil-small.en", 166.1e6)):
        shapes = wo.param_shapes(wo.PRESETS[name])
        n = sum(int(np.prod(s)) for k, s in shapes.items() if k != "model.encoder.embed_positions.weight")
        n += int(np.prod(shapes["model.encoder.embed_positions.weight"]))
        assert abs(n - expect) / expect < 2e-3, (name, n)


def test_config_validation():
    with pytest.raises(ValueError):
        WhisperB200Config.from_any({"d_model": 384, "encoder_attention_heads": 4, "decoder_attention_heads": 4})   # head_dim 96
    with pytest.raises(ValueError):
        WhisperB200Config.from_any({"dropout": 0.1})
    c = WhisperB200Config.from_any(wo.PRESETS["distil-small.en"].to_dict())
    assert c.d_model == 768 and c.decoder_layers == 4


def test_decay_split_follows_reference_rule():
    m = _tiny()
    names = get_parameter_names(m, [nn.LayerNorm])
    assert not any("layer_norm" in n for n in names)
    for p in m.model.encoder.parameters():
        p.requires_grad = False
    decay, no_decay = decay_split(m)
    by_id = {id(p): n for n, p in m.named_parameters()}
    assert all("bias" not in by_id[id(p)] and "layer_norm" not in by_id[id(p)] for p in decay)
    assert all(("bias" in by_id[id(p)]) or ("layer_norm" in by_id[id(p)]) for p in no_decay)
    assert all(by_id[id(p)].startswith("model.decoder") for p in decay + no_decay)
    assert any(by_id[id(p)].endswith("embed_tokens.weight") for p in decay)


def test_flat_buffers_rehome_params_and_grads():
    m = _tiny()
    decay, no_decay = decay_split(m)
    before = {n: p.detach().clone() for n, p in m.named_parameters()}
    fb = FlatBuffers([decay, no_decay])
    for n, p in m.named_parameters():
        assert torch.equal(p, before[n])
        if p.requires_grad:
            assert p.grad is not None and p.grad.data_ptr() >= fb.grad.data_ptr()
            assert p.data_ptr() % 16 == 0
    decay[0].grad.fill_(2.0)
    assert float(fb.grad.sum()) == 2.0 * decay[0].numel()
    assert fb.groups[0][1] + fb.groups[1][1] == fb.data.numel()
    assert m.proj_out.weight.data_ptr() == m.model.decoder.embed_tokens.weight.data_ptr()     # tie survives re-homing


def test_shift_tokens_right_matches_oracle():
    labels = torch.tensor([[-100, -100, 7, 8, -100], [5, 6, -100, -100, -100]])
    assert torch.equal(engine.shift_tokens_right(labels, 500, 501), wo.shift_tokens_right(labels, 500, 501))


def test_feature_extractor_host_side():
    fe = WhisperFeatureExtractorB200(80)
    np.testing.assert_allclose(fe.mel_filters, lo.mel_filter_bank(80), atol=1e-12)
    np.testing.assert_allclose(slaney_mel_filter_bank(201, 128, 0.0, 8000.0, 16000), lo.mel_filter_bank(128), atol=1e-12)
    clips = fe.pad_or_trim([np.ones(10, dtype=np.float32), np.ones(500000, dtype=np.float32)])
    assert clips.shape == (2, 480000) and clips[0, 10:].sum() == 0 and clips[1].sum() == 480000
    with pytest.raises(ValueError):
        fe([np.zeros(4, dtype=np.float32)], sampling_rate=8000)
    with pytest.raises(ValueError):
        WhisperFeatureExtractorB200(80, n_fft=512)


def test_sinusoids_closed_form():
    assert torch.allclose(sinusoids(1500, 1280), wo.sinusoids(1500, 1280))


def test_generate_rejects_unsupported_decoding_modes_before_touching_the_device():
    """generate() is greedy only (ref:training/run_distillation.py:1526 with its default gen_kwargs); anything else must
    raise instead of silently decoding differently -- checked here without a GPU because the argument screen comes first."""
    import pytest
    from distil_whisper_b200.modeling import DistilWhisperB200ForConditionalGeneration
    m = DistilWhisperB200ForConditionalGeneration(wo.PRESETS["tiny-student"].to_dict())
    feats = torch.zeros((1, 80, 100))
    for kw in (dict(num_beams=4), dict(do_sample=True), dict(forced_decoder_ids=[(1, 2)]), dict(prompt_ids=[1, 2]), dict(made_up_flag=1)):
        with pytest.raises(NotImplementedError):
            m.generate(feats, **kw)
    with pytest.raises(ValueError):                  # timestamps need the generation config's no_timestamps_token_id (as in HF)
        m.generate(feats, return_timestamps=True)


def test_pad_label_rows_matches_tokenizer_pad_semantics():
    """Host half of the device collator (distil_whisper_b200/data.py): right padding to max_target_length / longest, lengths."""
    import pytest
    from distil_whisper_b200.data import pad_label_rows
    rows = [[5, 6, 7, 8], [9], [1, 2]]
    t, l = pad_label_rows(rows, 99, 6, "max_length")
    assert t.tolist() == [[5, 6, 7, 8, 99, 99], [9, 99, 99, 99, 99, 99], [1, 2, 99, 99, 99, 99]] and l.tolist() == [4, 1, 2]
    t, l = pad_label_rows(rows, 99, None, "longest")
    assert t.shape == (3, 4) and t[1].tolist() == [9, 99, 99, 99]
    with pytest.raises(ValueError):
        pad_label_rows(rows, 99, 3, "max_length")
    with pytest.raises(ValueError):
        pad_label_rows(rows, 99, None, "max_length")


def test_save_pretrained_round_trips_and_is_hf_loadable(tmp_path):
    """Model: config.json + model.safetensors with HF names -> from_pretrained (ours) and transformers' own loader agree.
    Feature extractor: preprocessor_config.json loadable by transformers.WhisperFeatureExtractor (ref :1071, :1641, :1754)."""
    import pytest
    import torch
    from distil_whisper_b200.feature_extraction import WhisperFeatureExtractorB200
    from distil_whisper_b200.modeling import DistilWhisperB200ForConditionalGeneration
    from oracle import whisper_oracle as wo
    sc = wo.PRESETS["tiny-student"]
    m = DistilWhisperB200ForConditionalGeneration(sc.to_dict())
    m.load_hf_state_dict(wo.init_state_dict(sc, 3))
    m.generation_config = {"decoder_start_token_id": 501, "lang_to_id": {"<|en|>": 503}}
    d = tmp_path / "ckpt"
    m.save_pretrained(str(d))
    assert (d / "model.safetensors").exists() and (d / "config.json").exists()
    m2 = DistilWhisperB200ForConditionalGeneration.from_pretrained(str(d))
    for (k, a), (k2, b) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert k == k2 and torch.equal(a, b), k
    assert m2.generation_config["lang_to_id"] == {"<|en|>": 503}
    assert m2.proj_out.weight.data_ptr() == m2.model.decoder.embed_tokens.weight.data_ptr()
    transformers = pytest.importorskip("transformers")
    hf = transformers.WhisperForConditionalGeneration.from_pretrained(str(d))
    hsd = hf.state_dict()
    for k, v in m.state_dict().items():
        assert torch.equal(hsd[k], v), k
    m.save_pretrained(str(tmp_path / "bin"), safe_serialization=False)
    m3 = DistilWhisperB200ForConditionalGeneration.from_pretrained(str(tmp_path / "bin"), torch_dtype=torch.bfloat16)
    assert m3.model.decoder.layers[0].fc1.weight.dtype == torch.bfloat16
    fe = WhisperFeatureExtractorB200(128)
    fe.save_pretrained(str(d))
    hfe = transformers.WhisperFeatureExtractor.from_pretrained(str(d))
    assert hfe.feature_size == 128 and hfe.n_fft == 400 and hfe.hop_length == 160 and hfe.chunk_length == 30
    assert WhisperFeatureExtractorB200.from_pretrained(str(d)).feature_size == 128
    with pytest.warns(UserWarning):
        m.gradient_checkpointing_enable()
