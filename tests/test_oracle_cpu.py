"""CPU suite, part 1: the oracle against the golden vectors the reference produced
(oracle/validate_against_reference.py), and the plain-C delay oracle against the numpy one."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from conftest import GOLD, REPO
from dia_tts_prune_b200 import synthetic as SY
from dia_tts_prune_b200.config import DiaConfig, dia_1_6b_config, tiny_config
from oracle import delay_oracle, dia_oracle as O

DELAYS = [0, 8, 9, 10, 11, 12, 13, 14, 15]


# ---- integer delay path ---------------------------------------------------------------------------
def test_delay_known_answer(gold_delay):
    g = gold_delay
    x = np.array(g["input"], dtype=np.int32)
    ap = delay_oracle.apply_audio_delay(x, g["pad"], g["bos"], g["delay"])
    assert ap.tolist() == g["apply"] and ap.dtype == np.int32
    rv = delay_oracle.revert_audio_delay(ap, g["pad"], g["delay"], g["T"])
    assert rv.tolist() == g["revert"]
    # SURVEY.md Appendix C, verbatim
    assert ap[0, :, 1].tolist() == [1026, 21, 51, 81, 111, 141]
    assert ap[0, :, 2].tolist() == [1026, 1026, 32, 62, 92, 122]
    assert rv[0, :, 2].tolist() == [32, 62, 92, 122, 122, 122]
    x9 = np.array(g["input9"], dtype=np.int32)
    assert delay_oracle.apply_audio_delay(x9, 1025, 1026, DELAYS).tolist() == g["apply9"]
    assert delay_oracle.revert_audio_delay(x9, 1025, DELAYS, 24).tolist() == g["revert9"]


def test_delay_index_dtypes():
    t, i = delay_oracle.build_delay_indices(2, 5, 9, DELAYS)
    assert t.dtype == np.int32 and i.dtype == np.int64 and i.shape == (90, 3)
    t, i = delay_oracle.build_revert_indices(2, 5, 9, DELAYS)
    assert t.dtype == np.int64 and i.dtype == np.int64
    assert t.max() == 4


@pytest.fixture(scope="module")
def c_delay():
    subprocess.run(["make", "-C", os.path.join(REPO, "oracle")], check=True, capture_output=True)
    lib = ctypes.CDLL(os.path.join(REPO, "oracle", "_build", "libdelay_ref.so"))
    return lib


@pytest.mark.parametrize("B,T,C,dl", [(1, 1, 9, DELAYS), (1, 5, 9, DELAYS), (2, 16, 9, DELAYS), (3, 100, 9, DELAYS),
                                      (1, 3072, 9, DELAYS), (2, 17, 4, [3, 0, 7, 1]), (0, 4, 9, DELAYS)])
def test_c_delay_oracle_matches_numpy(c_delay, B, T, C, dl):
    rng = np.random.default_rng(B * 1000 + T)
    x = rng.integers(0, 1024, size=(B, T, C), dtype=np.int32)
    d = (ctypes.c_int32 * C)(*dl)
    out = np.empty_like(x)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)   # noqa: E731
    assert c_delay.delay_ref_apply_i32(p(x), p(out), B, T, C, d, 1025, 1026) == 0
    assert np.array_equal(out, delay_oracle.apply_audio_delay(x, 1025, 1026, dl))
    assert c_delay.delay_ref_revert_i32(p(x), p(out), B, T, C, d, 1025, T) == 0
    assert np.array_equal(out, delay_oracle.revert_audio_delay(x, 1025, dl, T))


def test_delay_properties_full_size():
    """Size-independent properties at the BASELINE size (T = 3072 + 16)."""
    rng = np.random.default_rng(1)
    T = 3088
    x = rng.integers(0, 1024, size=(1, T, 9), dtype=np.int32)
    ap = delay_oracle.apply_audio_delay(x, 1025, 1026, DELAYS)
    rv = delay_oracle.revert_audio_delay(ap, 1025, DELAYS, T)
    assert np.array_equal(rv[:, : T - 15], x[:, : T - 15])        # revert o apply = identity away from the tail
    for c, d in enumerate(DELAYS):
        assert (ap[0, :d, c] == 1026).all() and (ap[0, d:, c] == x[0, : T - d, c]).all()
    assert (ap != 1025).all()                                      # the PAD branch of apply is unreachable


# ---- sampling filter ----------------------------------------------------------------------------------
def test_sampling_known_answers(gold_sampling):
    assert len(gold_sampling["cases"]) >= 8
    for c in gold_sampling["cases"]:
        got = O.filtered_probs(torch.tensor(c["logits"]), c["temperature"], c["top_p"], c["top_k"])
        assert torch.allclose(got, torch.tensor(c["probs"]), atol=1e-6, rtol=0)
    p = O.filtered_probs(torch.log(torch.tensor([[.5, .3, .1, .05, .05]])), 1.0, 0.75, None)[0]
    assert torch.allclose(p, torch.tensor([.625, .375, 0, 0, 0]), atol=1e-6)
    p = O.filtered_probs(torch.log(torch.tensor([[.5, .3, .1, .05, .05]])), 1.0, 0.80, None)[0]
    assert (p > 0).sum() == 3                                      # strict '>' keeps the third entry
    p = O.filtered_probs(torch.log(torch.tensor([[.5, .3, .1, .05, .05]])), 1.0, 1.0, 4)[0]
    assert (p > 0).sum() == 5                                      # ties with the k-th value stay
    assert O.sample_next_token(torch.tensor([[1., 3., 3., 2.]]), 0.0, 0.95, 35).item() == 1


def test_cfg_masks():
    cfg = tiny_config()
    lg = torch.zeros(2, 9, 1028)
    lg[1] = 1.0
    g = O.cfg_combine_and_mask(cfg, lg, 3.0)
    assert torch.isinf(g[1:, 1024]).all() and not torch.isinf(g[0, 1024])
    assert torch.isinf(g[:, 1025]).all() and torch.isinf(g[:, 1026]).all()
    assert not torch.isinf(g[:, 1027]).any()                       # Q5: 1027 stays legal
    assert g[0, 0].item() == 4.0                                   # Q4: cond + 3 (cond - uncond)


# ---- model path ------------------------------------------------------------------------------------------
def test_param_order_and_shapes():
    import json
    with open(os.path.join(GOLD, "param_names_dia16b.json")) as f:
        g = json.load(f)
    cfg = dia_1_6b_config()
    assert O.param_names(cfg) == g["names"]
    assert {k: list(v) for k, v in O.param_shapes(cfg).items()} == g["shapes"]
    n = sum(int(np.prod(s)) for s in g["shapes"].values())
    assert abs(n - 1_611.2e6) < 1e5                                # "1.6B"


@pytest.fixture(scope="module")
def tiny_sd(gold_tiny):
    cfg = DiaConfig.model_validate_json(str(gold_tiny["config_json"]))
    assert cfg == tiny_config()
    sd = SY.synthetic_state_dict(O.param_shapes(cfg), int(gold_tiny["weight_seed"]))
    return cfg, sd


def test_weight_recipe_reproduces_golden_fingerprint(tiny_sd, gold_tiny):
    cfg, sd = tiny_sd
    assert SY.weights_fingerprint(sd) == str(gold_tiny["fingerprint"]), \
        "torch's CPU normal_ stream differs on this machine from the one the golden vectors were made on"


def test_oracle_reproduces_tiny_golden(tiny_sd, gold_tiny):
    cfg, sd = tiny_sd
    steps = [int(s) for s in gold_tiny["logits_steps"]]
    tr = O.generate(sd, cfg, str(gold_tiny["text"]), max_tokens=40, temperature=0.0, keep_logits_at=set(steps),
                    dead_cross_kv=False)
    assert torch.equal(tr.codes, torch.from_numpy(gold_tiny["codes"]))
    assert torch.equal(tr.grid, torch.from_numpy(gold_tiny["grid"]))
    for i, s in enumerate(steps):
        assert np.abs(tr.logits[s].numpy() - gold_tiny["logits"][i]).max() < 1e-5
    # the dead cross K/V projection changes nothing
    tr2 = O.generate(sd, cfg, str(gold_tiny["text"]), max_tokens=12, temperature=0.0, dead_cross_kv=True)
    assert torch.equal(tr2.codes, tr.codes[: tr2.codes.shape[0]])


def test_oracle_reproduces_tiny_clone_golden(tiny_sd, gold_tiny):
    cfg, sd = tiny_sd
    prompt = torch.from_numpy(gold_tiny["clone_prompt"])
    tr = O.generate(sd, cfg, str(gold_tiny["clone_text"]), max_tokens=int(gold_tiny["clone_max_tokens"]),
                    temperature=0.0, audio_prompt=prompt, audio_prompt_text=str(gold_tiny["clone_prompt_text"]),
                    dead_cross_kv=False)
    assert tr.prefill_step == int(gold_tiny["clone_prefill_step"]) == 21
    assert torch.equal(tr.codes, torch.from_numpy(gold_tiny["clone_codes"]))
    # Q3: the delay tail is PAD, so the first 14 predictions after a prompt are discarded
    assert (tr.codes[:14, 0] == 1025).all()
    with pytest.raises(ValueError):
        O.generate(sd, cfg, "x", audio_prompt=prompt)


def test_oracle_full_size_two_steps_against_golden(gold_full):
    """Dia-1.6B: the oracle's logits at steps 1-2 against the reference-generated fixture."""
    cfg = dia_1_6b_config()
    sd = SY.synthetic_state_dict(O.param_shapes(cfg), int(gold_full["weight_seed"]))
    sub = [n for n in sd if "layers.0." in n or "logits" in n]
    assert SY.weights_fingerprint(sd, sub) == str(gold_full["fingerprint"])
    tr = O.generate(sd, cfg, str(gold_full["text"]), max_tokens=3, temperature=0.0, keep_logits_at={1, 2},
                    dead_cross_kv=False)
    steps = gold_full["logits_steps"].tolist()
    for s in (1, 2):
        assert np.abs(tr.logits[s].numpy() - gold_full["logits"][steps.index(s)]).max() < 1e-4
    assert torch.equal(tr.codes, torch.from_numpy(gold_full["codes"][: tr.codes.shape[0]]))
    assert gold_full["margins"].min() > 3e-4                       # the seed screen


def test_effective_text_and_text_encoding():
    cfg = tiny_config()
    assert O.effective_text("[S1] a. [S2] b.", None).endswith(" [S1]")
    assert O.effective_text("[S1] a.", None).endswith(" [S2]")
    assert O.effective_text("hello", None) == "hello [S2]"
    assert O.effective_text("[S2] b", "[S1] p") == "[S1] p [S2] b [S1]"
    t = O.encode_text(cfg, "[S1] hi [S2]")
    assert t.shape == (1, cfg.data.text_length) and t[0, 0] == 1 and t[0, 5] == 2 and t[0, 6] == 0
