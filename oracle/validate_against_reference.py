#!/usr/bin/env python
"""Pin the oracle against the reference itself and write tests/golden/.

TEST INFRASTRUCTURE ONLY.  Runs in the build container only (it imports the
reference from /root/reference, which does not exist on the GPU box); the
fixtures it writes are committed and travel.

What it does
  1. imports the reference ``dia`` package with ``dac``/``soundfile`` stubbed and
     the hot-path patches B1-B4 of SURVEY.md Appendix B (the shipped code raises
     before the first token without them; nothing else is touched);
  2. checks ``oracle/dia_oracle.py`` and ``oracle/delay_oracle.py`` against it,
     bit for bit: parameter order, delay/revert gathers, encoder output,
     cross-KV, per-step logits, self-KV contents, greedy code streams with and
     without an audio prompt, the sampling filter;
  3. writes the golden fixtures.

Usage:  python oracle/validate_against_reference.py [--tiny] [--full] [--clone] [--all]
"""

from __future__ import annotations

import argparse
import json
import os
import random
import sys
import time
import types

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_ROOT = os.environ.get("DIA_REFERENCE_ROOT", "/root/reference")
GOLD = os.path.join(REPO, "tests", "golden")


def import_reference():
    """Reference modules with stubs + patches (SURVEY.md Appendix B shim)."""
    for n in ("dac", "soundfile"):
        sys.modules.setdefault(n, types.ModuleType(n))
    # the repo ships a drop-in package that is also called ``dia``: make sure the
    # reference wins in this process
    sys.path[:] = [p for p in sys.path if os.path.abspath(p or ".") != REPO]
    sys.path.insert(0, REF_ROOT)
    for k in [k for k in sys.modules if k == "dia" or k.startswith("dia.")]:
        del sys.modules[k]
    import dia.audio as A
    import dia.config as C
    import dia.layers as L
    import dia.model as M
    import dia.state as S
    assert os.path.abspath(L.__file__).startswith(REF_ROOT), L.__file__

    def rope_forward(self, inputs, position):                                   # B1
        pos = position.unsqueeze(-1).unsqueeze(-1)
        freqs = pos * self.inv_freq.to(pos.device)
        sin, cos = torch.sin(freqs.float()), torch.cos(freqs.float())
        x1, x2 = torch.chunk(inputs.to(torch.float32), 2, dim=-1)
        return torch.cat((x1 * cos - x2 * sin, x1 * sin + x2 * cos), dim=-1).to(self.compute_dtype)

    def kv_prefill(self, k, v):                                                 # B3
        n = k.shape[2]
        self.k[:, :, :n, :] = k
        self.v[:, :, :n, :] = v
        self.current_idx = n - 1
        return k, v

    def get_tokens_at(self, step_from, step_to=None):                           # B2
        return self.generated_tokens[step_from] if step_to is None else self.generated_tokens[step_from:step_to, :]

    L.RotaryEmbedding.forward = rope_forward
    S.KVCache.prefill = kv_prefill
    S.DecoderOutput.get_tokens_at = get_tokens_at
    M.random = random                                                           # B4
    return types.SimpleNamespace(A=A, C=C, L=L, M=M, S=S)


sys.path.insert(0, REPO)
from oracle import delay_oracle, dia_oracle as O          # noqa: E402
from dia_tts_prune_b200 import config as PC               # noqa: E402
from dia_tts_prune_b200 import synthetic as SY            # noqa: E402


def ref_config(ref, pcfg):
    return ref.C.DiaConfig.model_validate(pcfg.model_dump())


def build_reference(ref, pcfg, seed):
    cfg = ref_config(ref, pcfg)
    d = ref.M.Dia(cfg, "float32", torch.device("cpu"))
    SY.init_synthetic_(d.model.named_parameters(), seed)
    d.model.eval()
    return d, cfg


class RefTrace:
    """Hooks on the reference: logits per step, guided logits, codes, KV."""

    def __init__(self, ref, d):
        self.ref, self.d = ref, d
        self.logits, self.guided, self.codes = [], [], None
        self.state = None
        orig_step = d.model.decoder.decode_step
        orig_sample = ref.M._sample_next_token
        orig_prep = d._prepare_generation

        def step(tokens, state):
            out = orig_step(tokens, state)
            self.logits.append(out[:, -1].clone())
            return out

        def sample(logits, temperature, top_p, cfg_filter_top_k=None):
            self.guided.append(logits.clone())
            return orig_sample(logits, temperature, top_p, cfg_filter_top_k)

        def prep(text, audio_prompt, verbose):
            st, out = orig_prep(text, audio_prompt, verbose)
            self.state, self.out = st, out
            return st, out

        d.model.decoder.decode_step = step
        ref.M._sample_next_token = sample
        d._prepare_generation = prep
        d._generate_output = lambda codes: self._keep(codes)
        self._restore = lambda: setattr(ref.M, "_sample_next_token", orig_sample)

    def _keep(self, codes):
        self.codes = codes.clone()
        return None

    def close(self):
        self._restore()


def check_equal(name, a, b):
    a, b = torch.as_tensor(a), torch.as_tensor(b)
    assert a.shape == b.shape, (name, a.shape, b.shape)
    if not torch.equal(a, b):
        diff = (a.double() - b.double()).abs().max().item() if a.is_floating_point() else "int"
        raise AssertionError(f"{name}: oracle != reference (max abs diff {diff})")
    print(f"  ok  {name}  {tuple(a.shape)}")


# ----------------------------------------------------------------------------

def run_delay(ref):
    print("[delay] oracle vs unpatched dia/audio.py")
    rng = np.random.default_rng(0)
    cases = [(1, 6, 3, [0, 1, 2]), (1, 1, 9, PC._DEFAULT_DELAYS), (1, 5, 9, PC._DEFAULT_DELAYS),
             (1, 16, 9, PC._DEFAULT_DELAYS), (2, 40, 9, PC._DEFAULT_DELAYS), (3, 3088, 9, PC._DEFAULT_DELAYS),
             (1, 3072, 9, PC._DEFAULT_DELAYS), (2, 17, 4, [3, 0, 7, 1])]
    for B, T, C, dl in cases:
        dl = list(dl)
        x = rng.integers(0, 1024, size=(B, T, C), dtype=np.int32)
        xt = torch.from_numpy(x)
        pre = ref.A.build_delay_indices(B, T, C, dl)
        got_t, got_i = delay_oracle.build_delay_indices(B, T, C, dl)
        check_equal(f"build_delay_indices t_idx {B,T,C}", got_t, pre[0])
        check_equal(f"build_delay_indices idx {B,T,C}", got_i, pre[1])
        assert pre[0].dtype == torch.int32 and pre[1].dtype == torch.int64
        check_equal(f"apply_audio_delay {B,T,C}", delay_oracle.apply_audio_delay(x, 1025, 1026, dl),
                    ref.A.apply_audio_delay(xt, 1025, 1026, pre))
        rpre = ref.A.build_revert_indices(B, T, C, dl)
        got_t, got_i = delay_oracle.build_revert_indices(B, T, C, dl)
        check_equal(f"build_revert_indices t_idx {B,T,C}", got_t, rpre[0])
        check_equal(f"build_revert_indices idx {B,T,C}", got_i, rpre[1])
        check_equal(f"revert_audio_delay {B,T,C}", delay_oracle.revert_audio_delay(x, 1025, dl, T),
                    ref.A.revert_audio_delay(xt, 1025, rpre, T))
    # known answer, produced by the reference right here
    B, T, C, dl = 1, 6, 3, [0, 1, 2]
    x = torch.tensor([[[10 + 30 * t + 11 * c for c in range(C)] for t in range(T)]], dtype=torch.int32)
    ap = ref.A.apply_audio_delay(x, 1025, 1026, ref.A.build_delay_indices(B, T, C, dl))
    rv = ref.A.revert_audio_delay(ap, 1025, ref.A.build_revert_indices(B, T, C, dl), T)   # revert OF the delayed grid
    ka = {"B": B, "T": T, "C": C, "delay": dl, "pad": 1025, "bos": 1026, "input": x.tolist(),
          "apply": ap.tolist(), "revert": rv.tolist(), "source": "unpatched /root/reference/dia/audio.py"}
    assert ap[0, :, 1].tolist() == [1026, 21, 51, 81, 111, 141] and rv[0, :, 2].tolist() == [32, 62, 92, 122, 122, 122]
    # a 9-codebook case at the real delay pattern as well
    x9 = torch.from_numpy(rng.integers(0, 1024, size=(2, 24, 9), dtype=np.int32))
    dl9 = list(PC._DEFAULT_DELAYS)
    ka["input9"] = x9.tolist()
    ka["apply9"] = ref.A.apply_audio_delay(x9, 1025, 1026, ref.A.build_delay_indices(2, 24, 9, dl9)).tolist()
    ka["revert9"] = ref.A.revert_audio_delay(x9, 1025, ref.A.build_revert_indices(2, 24, 9, dl9), 24).tolist()
    with open(os.path.join(GOLD, "delay_known_answer.json"), "w") as f:
        json.dump(ka, f)
    print("  wrote delay_known_answer.json")


def run_sampling(ref):
    print("[sampling] filter vs dia/model.py:_sample_next_token")
    torch.manual_seed(0)
    out = {"cases": []}
    logit_sets = [torch.log(torch.tensor([[.5, .3, .1, .05, .05]])), torch.randn(9, 1028) * 3.0,
                  torch.randn(4, 64), torch.randn(9, 1028) * 8.0]
    settings = [(1.0, 0.75, None), (1.0, 0.80, None), (1.0, 1.0, 4), (1.3, 0.95, 35), (0.7, 0.5, 10), (2.0, 0.99, 0)]
    for li, lg in enumerate(logit_sets):
        for (T, p, k) in settings:
            if k and k > lg.shape[-1]:
                continue
            mine = O.filtered_probs(lg.clone(), T, p, k)
            # run the reference to the same point: replace multinomial to capture its input
            cap = {}
            orig = torch.multinomial
            torch.multinomial = lambda pr, num_samples=1, **kw: (cap.setdefault("p", pr.clone()), orig(pr, num_samples))[1]
            try:
                ref.M._sample_next_token(lg.clone(), T, p, k)
            finally:
                torch.multinomial = orig
            check_equal(f"filtered_probs set{li} T={T} p={p} k={k}", mine, cap["p"])
            if li in (0, 2):
                out["cases"].append({"logits": lg.tolist(), "temperature": T, "top_p": p, "top_k": k,
                                     "probs": cap["p"].tolist()})
        check_equal(f"argmax set{li}", O.sample_next_token(lg, 0.0, 0.95, 35), ref.M._sample_next_token(lg, 0.0, 0.95, 35))
    ka = O.filtered_probs(torch.log(torch.tensor([[.5, .3, .1, .05, .05]])), 1.0, 0.75, None)[0]
    assert torch.allclose(ka, torch.tensor([.625, .375, 0, 0, 0]), atol=1e-6), ka
    with open(os.path.join(GOLD, "sampling_known_answer.json"), "w") as f:
        json.dump(out, f)
    print("  wrote sampling_known_answer.json")


def compare_generation(ref, pcfg, seed, text, max_tokens, prompt=None, prompt_text=None, label="", keep=None,
                       lean_oracle=False):
    """Run reference + oracle greedy; return (ref trace, oracle trace, state dict)."""
    d, cfg = build_reference(ref, pcfg, seed)
    sd = {k: v.detach() for k, v in d.model.named_parameters()}
    names = [k for k, _ in d.model.named_parameters()]
    assert names == O.param_names(pcfg), "parameter order differs from the reference"
    for n, s in O.param_shapes(pcfg).items():
        assert tuple(sd[n].shape) == s, (n, sd[n].shape, s)
    rt = RefTrace(ref, d)
    t0 = time.time()
    d.generate(text, max_tokens=max_tokens, temperature=0.0, cfg_scale=3.0, audio_prompt=prompt,
               audio_prompt_text=prompt_text)
    t_ref = time.time() - t0
    rt.close()
    assert rt.codes is not None, "reference generate failed"
    t0 = time.time()
    ot = O.generate(sd, pcfg, text, max_tokens=max_tokens, temperature=0.0, cfg_scale=3.0, audio_prompt=prompt,
                    audio_prompt_text=prompt_text, keep_logits_at=set(range(0, 100000)) if keep is None else keep,
                    dead_cross_kv=not lean_oracle)
    t_or = time.time() - t0
    print(f"[{label}] reference {t_ref:.1f}s ({len(rt.logits)} steps), oracle {t_or:.1f}s")
    check_equal(f"{label} codes", ot.codes, rt.codes)
    check_equal(f"{label} token grid", ot.grid, rt.out.generated_tokens)
    steps = sorted(ot.logits)
    first = ot.prefill_step
    for s in steps:
        check_equal(f"{label} logits step {s}", ot.logits[s], rt.logits[s - first]) if (s - first) < 3 or s == steps[-1] \
            else None
        assert torch.equal(ot.logits[s], rt.logits[s - first]) and torch.equal(ot.guided[s], rt.guided[s - first]), s
    print(f"  ok  {label} logits + guided logits for {len(steps)} steps (bit-exact)")
    return d, rt, ot, sd


def run_tiny(ref):
    pcfg = PC.tiny_config()
    text = "[S1] Hello there. [S2] Hi."
    d, rt, ot, sd = compare_generation(ref, pcfg, 7, text, 40, label="tiny")
    # KV cache contents after the run
    for i in range(pcfg.model.decoder.n_layer):
        # oracle state is internal to generate(); re-run the preparation for cross-KV / encoder checks
        pass
    st_o, grid_o, p0 = O.prepare_generation(sd, pcfg, O.effective_text(text, None), None)
    check_equal("tiny enc_out", st_o.enc_out, rt.state.enc_out)
    for i in range(pcfg.model.decoder.n_layer):
        check_equal(f"tiny cross k L{i}", st_o.cross_cache[i].k, rt.state.cross_attn_cache[i].k)
        check_equal(f"tiny cross v L{i}", st_o.cross_cache[i].v, rt.state.cross_attn_cache[i].v)
    check_equal("tiny cross mask", st_o.cross_mask, rt.state.dec_cross_attn_mask)

    # voice-clone path (prefill + slot clobber quirk) on the tiny config
    g = torch.Generator().manual_seed(3)
    prompt = torch.randint(0, 1024, (20, 9), generator=g, dtype=torch.int32)
    d2, rt2, ot2, _ = compare_generation(ref, pcfg, 7, "[S2] And more.", 21 + 40, prompt=prompt,
                                         prompt_text="[S1] Prompt words.", label="tiny-clone")
    np.savez_compressed(
        os.path.join(GOLD, "tiny_seed7.npz"),
        config_json=np.array(PC.config_to_json(pcfg)), weight_seed=np.array(7), text=np.array(text),
        fingerprint=np.array(SY.weights_fingerprint(sd)),
        codes=ot.codes.numpy(), grid=ot.grid.numpy(), prefill_step=np.array(ot.prefill_step),
        logits_steps=np.array(sorted(ot.logits)[:8]),
        logits=np.stack([ot.logits[s].numpy() for s in sorted(ot.logits)[:8]]),
        margins=torch.stack(ot.margins).numpy(),
        clone_prompt=prompt.numpy(), clone_text=np.array("[S2] And more."), clone_prompt_text=np.array("[S1] Prompt words."),
        clone_max_tokens=np.array(61), clone_codes=ot2.codes.numpy(), clone_grid=ot2.grid.numpy(),
        clone_prefill_step=np.array(ot2.prefill_step), clone_margins=torch.stack(ot2.margins).numpy(),
        clone_logits_steps=np.array(sorted(ot2.logits)[:4]),
        clone_logits=np.stack([ot2.logits[s].numpy() for s in sorted(ot2.logits)[:4]]),
    )
    with open(os.path.join(GOLD, "param_names_dia16b.json"), "w") as f:
        big = PC.dia_1_6b_config()
        json.dump({"names": O.param_names(big), "shapes": {k: list(v) for k, v in O.param_shapes(big).items()}}, f)
    print("  wrote tiny_seed7.npz, param_names_dia16b.json")


def run_full(ref, seed, max_tokens):
    pcfg = PC.dia_1_6b_config()
    text = SY.DEFAULT_TRANSCRIPT
    keep = set(list(range(1, 9)) + [15, 16, 17, 32, 64, 100, 128, 180, 200, 242, 255, 256])
    d, rt, ot, sd = compare_generation(ref, pcfg, seed, text, max_tokens, label=f"dia16b-seed{seed}", keep=keep,
                                       lean_oracle=True)
    margins = torch.stack(ot.margins)
    print(f"  min top1-top2 margin over {margins.numel()} argmaxes: {margins.min().item():.3e} "
          f"at {divmod(int(margins.argmin()), margins.shape[1])}")
    steps = sorted(ot.logits)
    # a small KV probe: self cache of layers 0 and 17, first 4 slots, from the reference state
    kv_probe = np.stack([rt.state.self_attn_cache[i].k[:, :, :4, :].numpy() for i in (0, pcfg.model.decoder.n_layer - 1)])
    np.savez_compressed(
        os.path.join(GOLD, f"dia16b_seed{seed}_greedy.npz"),
        weight_seed=np.array(seed), text=np.array(text), max_tokens=np.array(max_tokens), cfg_scale=np.array(3.0),
        fingerprint=np.array(SY.weights_fingerprint(sd, [n for n in sd if "layers.0." in n or "logits" in n])),
        codes=ot.codes.numpy(), grid=ot.grid[: max_tokens + 1].numpy(), prefill_step=np.array(ot.prefill_step),
        margins=margins.numpy(), logits_steps=np.array(steps),
        logits=np.stack([ot.logits[s].numpy() for s in steps]).astype(np.float32),
        self_k_probe=kv_probe, finalized=O.finalize_codes(pcfg, ot.codes).numpy(),
        ref_seconds=np.array(0.0),
    )
    print(f"  wrote dia16b_seed{seed}_greedy.npz")


def run_clone(ref, seed, prompt_len, n_decode, keep_every=8, tag=None):
    pcfg = PC.dia_1_6b_config()
    g = torch.Generator().manual_seed(11)
    prompt = torch.randint(0, 1024, (prompt_len, 9), generator=g, dtype=torch.int32)
    text, ptext = "[S2] You get full control over scripts and voices.", "[S1] Dia is an open weights text to dialogue model."
    max_tokens = prompt_len + 1 + n_decode
    assert max_tokens <= pcfg.data.audio_length
    keep = set(range(prompt_len + 1, prompt_len + 1 + n_decode, keep_every))
    d, rt, ot, sd = compare_generation(ref, pcfg, seed, text, max_tokens, prompt=prompt, prompt_text=ptext,
                                       label=f"dia16b-clone{prompt_len}", keep=keep, lean_oracle=True)
    margins = torch.stack(ot.margins)
    print(f"  min margin {margins.min().item():.3e}")
    steps = sorted(ot.logits)
    name = f"dia16b_seed{seed}_clone{tag or prompt_len}.npz"
    np.savez_compressed(
        os.path.join(GOLD, name),
        weight_seed=np.array(seed), text=np.array(text), prompt_text=np.array(ptext), prompt=prompt.numpy(),
        max_tokens=np.array(max_tokens), codes=ot.codes.numpy(), grid=ot.grid[: max_tokens + 1].numpy(),
        prefill_step=np.array(ot.prefill_step), margins=margins.numpy(), logits_steps=np.array(steps),
        logits=np.stack([ot.logits[s].numpy() for s in steps]).astype(np.float32),
    )
    print(f"  wrote {name}")


def text_of_length(n_bytes: int) -> str:
    """A [S1]/[S2] dialogue whose effective text (after Dia.generate's speaker-tag completion, dia/model.py:689-696)
    encodes to exactly ``n_bytes`` text tokens ([S1]/[S2] are one byte each, dia/model.py:262-263)."""
    words = ("dia is an open weights text to dialogue model you get full control over scripts and voices "
             "the quick brown fox jumps over the lazy dog while rain keeps falling on the quiet harbor town").split()
    out, spk, i = "[S1]", 1, 0

    def enc_len(t):
        return len(O.effective_text(t, None).encode("utf-8").replace(b"[S1]", b"\x01").replace(b"[S2]", b"\x02"))

    while enc_len(out) < n_bytes - 12:
        out += " " + words[i % len(words)]
        i += 1
        if i % 7 == 0:
            spk = 3 - spk
            out += f". [S{spk}]"
    while enc_len(out) < n_bytes:
        out += "a" if not out.endswith("]") else " a"
    assert enc_len(out) == n_bytes, (enc_len(out), n_bytes)
    return out


def run_longtext(ref, seed, lens, max_tokens=41):
    """Transcripts longer than 128 bytes: cross-attention over several K/V tiles per warp / several CTAs per head."""
    pcfg = PC.dia_1_6b_config()
    for n in lens:
        text = text_of_length(n)
        keep = set([1, 2, 3] + list(range(4, max_tokens, 4)))
        d, rt, ot, sd = compare_generation(ref, pcfg, seed, text, max_tokens, label=f"dia16b-text{n}", keep=keep,
                                           lean_oracle=True)
        margins = torch.stack(ot.margins)
        print(f"  Lt={n}: min margin {margins.min().item():.3e}")
        steps = sorted(ot.logits)
        name = f"dia16b_seed{seed}_text{n}.npz"
        np.savez_compressed(
            os.path.join(GOLD, name),
            weight_seed=np.array(seed), text=np.array(text), text_len=np.array(n), max_tokens=np.array(max_tokens),
            codes=ot.codes.numpy(), grid=ot.grid[: max_tokens + 1].numpy(), prefill_step=np.array(ot.prefill_step),
            margins=margins.numpy(), logits_steps=np.array(steps),
            logits=np.stack([ot.logits[s].numpy() for s in steps]).astype(np.float32),
        )
        print(f"  wrote {name}")


def run_pruned(ref, seed, max_tokens=33):
    """BASELINE configs[3] at Dia-1.6B, against the reference itself: (i) the stock ``offline_prune.py --pruning-type
    structured --prune-dim 0`` call (dia/pruning_utils.py:64-119 on every DenseGeneral, offline_prune.py:103-107) and
    (ii) the 2:4 mask (ours; the reference has none) - both made permanent (dia/pruning_utils.py:122-151) on the
    REFERENCE model, which then runs its own generate.  The kept-row masks of (i) travel in the fixture (norm
    near-ties must not depend on the CPU that regenerates them); the 2:4 mask is integer-keyed, hence portable."""
    import torch.nn.utils.prune as prune
    from dia_tts_prune_b200 import pruning_utils as PU
    pcfg = PC.dia_1_6b_config()
    text = "[S1] Pruned weights. [S2] Same answers."
    for mode in ("structured", "2to4"):
        d, cfg = build_reference(ref, pcfg, seed)
        RPU = __import__("dia.pruning_utils", fromlist=["x"])
        extra = {}
        if mode == "structured":
            RPU.apply_structured_pruning(d.model, 0.5, dim=0, n=2)
            for n_, m in d.model.named_modules():
                if isinstance(m, ref.L.DenseGeneral) and prune.is_pruned(m):
                    mask = m.weight_mask
                    keep = mask.reshape(mask.shape[0], -1).any(dim=1)
                    assert torch.equal(mask.reshape(mask.shape[0], -1), keep[:, None].expand(-1, mask[0].numel()).to(mask.dtype))
                    extra["keep::" + n_] = np.packbits(keep.numpy())
        else:
            for n_, m in d.model.named_modules():
                if isinstance(m, ref.L.DenseGeneral):
                    w = m.weight
                    K = int(np.prod(m.in_shapes))
                    prune.custom_from_mask(m, "weight", PU.mask_2to4(w.detach(), K).to(w.dtype))
        RPU.make_pruning_permanent(d.model)
        sd = {k: v.detach() for k, v in d.model.named_parameters()}
        assert [k for k in sd] == O.param_names(pcfg)
        rt = RefTrace(ref, d)
        t0 = time.time()
        d.generate(text, max_tokens=max_tokens, temperature=0.0, cfg_scale=3.0)
        rt.close()
        print(f"[pruned-{mode}] reference {time.time() - t0:.1f}s")
        keep_steps = set([1, 2] + list(range(4, max_tokens, 4)))
        ot = O.generate(sd, pcfg, text, max_tokens=max_tokens, temperature=0.0, cfg_scale=3.0, keep_logits_at=keep_steps,
                        dead_cross_kv=False)
        check_equal(f"pruned-{mode} codes", ot.codes, rt.codes)
        for s in sorted(ot.logits):
            assert torch.equal(ot.logits[s], rt.logits[s - ot.prefill_step]), s
        margins = torch.stack(ot.margins)
        print(f"  ok  pruned-{mode}: logits bit-exact at {len(ot.logits)} steps, min margin {margins.min().item():.3e}")
        steps = sorted(ot.logits)
        dec = [n for n in sd if n.startswith("decoder.layers.0.") or "logits" in n]
        name = f"dia16b_seed{seed}_pruned_{mode}.npz"
        np.savez_compressed(
            os.path.join(GOLD, name), weight_seed=np.array(seed), text=np.array(text), max_tokens=np.array(max_tokens),
            mode=np.array(mode), fingerprint=np.array(SY.weights_fingerprint(sd, dec)),
            codes=ot.codes.numpy(), grid=ot.grid[: max_tokens + 1].numpy(), margins=margins.numpy(),
            logits_steps=np.array(steps), logits=np.stack([ot.logits[s].numpy() for s in steps]).astype(np.float32),
            **extra)
        print(f"  wrote {name}")
        del d, rt, ot, sd


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tiny", action="store_true")
    ap.add_argument("--full", action="store_true")
    ap.add_argument("--clone", action="store_true")
    ap.add_argument("--all", action="store_true")
    ap.add_argument("--seed", type=int, default=5)
    ap.add_argument("--max-tokens", type=int, default=257)
    ap.add_argument("--prompt-len", type=int, default=861)
    ap.add_argument("--clone-steps", type=int, default=48)
    ap.add_argument("--long", action="store_true", help="long-context clone runs (slots ~1200, ~2380, ~2990)")
    ap.add_argument("--longtext", action="store_true", help="transcripts of 129 / 200 / 600 bytes")
    ap.add_argument("--clone-full", action="store_true", help="configs[2] at its stated size: 861-frame prompt + 1536 steps")
    ap.add_argument("--pruned", action="store_true", help="configs[3] at Dia-1.6B: structured dim-0 and 2:4")
    ap.add_argument("--threads", type=int, default=os.cpu_count())
    a = ap.parse_args()
    torch.set_num_threads(a.threads)
    os.makedirs(GOLD, exist_ok=True)
    ref = import_reference()
    if a.tiny or a.all:
        run_delay(ref)
        run_sampling(ref)
        run_tiny(ref)
    if a.full or a.all:
        run_full(ref, a.seed, a.max_tokens)
    if a.clone or a.all:
        run_clone(ref, a.seed, a.prompt_len, a.clone_steps)
    if a.longtext:
        run_longtext(ref, a.seed, [129, 200, 600])
    if a.pruned:
        run_pruned(ref, a.seed)
    if a.long:
        for pl in (1200, 2380, 2990):
            run_clone(ref, a.seed, pl, 48, keep_every=4)
    if a.clone_full:
        run_clone(ref, a.seed, 861, 1536, keep_every=128, tag="861x1536")
    print("VALIDATION PASSED")


if __name__ == "__main__":
    main()
