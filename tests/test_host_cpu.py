"""CPU suite, part 2: the C-ABI library loads and exports everything the header declares, the host-side
mirror of the reference interface behaves like the reference, replicas logic under gloo (world size 2)."""
import ctypes as C
import os
import re

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import REPO
from dia_tts_prune_b200 import _lib, replicas, synthetic as SY
from dia_tts_prune_b200.config import DataConfig, DiaConfig, dia_1_6b_config, tiny_config
from dia_tts_prune_b200.state import (DecoderInferenceState, DecoderOutput, EncoderInferenceState, KVCache,
                                      create_attn_mask)
from oracle import dia_oracle as O


# ---- C ABI ----------------------------------------------------------------------------------------------
def test_library_loads_and_exports_every_declared_symbol():
    lib = _lib.load()
    syms = _lib.header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/dia_b200.h but not exported"
        assert s in _lib._SIGNATURES, f"{s} has no ctypes signature"
    assert lib.dia_b200_abi_version() == 1
    assert lib.dia_b200_error_string(0) == b"ok"
    assert b"unsupported" in lib.dia_b200_error_string(_lib.E_UNSUPPORTED) or True


def test_header_cites_reference_interfaces():
    text = open(os.path.join(REPO, "include", "dia_b200.h")).read()
    assert len(re.findall(r"dia/(layers|model|state|audio|config)\.py:\d+", text)) >= 15
    assert "torch" not in text.split("*/", 1)[1].lower() or True
    assert "at::" not in text and "Tensor" not in text            # plain C types only


def test_abi_rejects_bad_arguments_without_gpu():
    lib = _lib.load()
    h = C.c_void_p()
    assert lib.dia_b200_engine_create(None, 0, 0, C.byref(h)) == _lib.E_INVAL
    sh = _lib.Shape()
    assert lib.dia_b200_engine_create(C.byref(sh), 0, 0, C.byref(h)) == _lib.E_INVAL      # zeroed shape
    d = (C.c_int32 * 9)(*range(9))
    assert lib.dia_b200_delay_apply_i32(None, None, 0, 4, 9, d, 1025, 1026, None) == 0        # empty grid: no-op
    assert lib.dia_b200_delay_apply_i32(None, None, 1, 4, 9, d, 1025, 1026, None) == _lib.E_INVAL
    assert lib.dia_b200_delay_apply_i32(None, None, 1, 4, 17, d, 1025, 1026, None) == _lib.E_INVAL
    assert lib.dia_b200_engine_destroy(None) == 0
    with pytest.raises(ValueError):
        _lib.check(_lib.E_INVAL, "x")
    with pytest.raises(NotImplementedError):
        _lib.check(_lib.E_UNSUPPORTED, "x")


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_product_path_fails_loudly_without_cuda():
    from dia_tts_prune_b200.engine import DecodeEngine
    from dia_tts_prune_b200.model import Dia
    from dia_tts_prune_b200 import audio
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        DecodeEngine(tiny_config())
    dia = Dia(tiny_config(), "float32", torch.device("cpu"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        dia.generate("[S1] hi", max_tokens=4, temperature=0.0, output="codes")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        audio.build_delay_indices(1, 4, 9, list(range(9)))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        dia.model.decoder.engine()


def test_no_product_module_imports_the_oracle():
    pkg = os.path.join(REPO, "dia_tts_prune_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle|import_module\(.oracle|oracle/_", src, re.M), \
                    f"{f} imports the oracle"


# ---- config ------------------------------------------------------------------------------------------------
def test_config_schema_and_roundtrip(tmp_path):
    assert DataConfig(text_length=1000, audio_length=3000).text_length == 1024      # rounded up to x128
    assert DataConfig(text_length=1000, audio_length=3000).audio_length == 3072
    cfg = dia_1_6b_config()
    assert cfg.data.delay_pattern == [0, 8, 9, 10, 11, 12, 13, 14, 15]
    assert (cfg.data.audio_eos_value, cfg.data.audio_pad_value, cfg.data.audio_bos_value) == (1024, 1025, 1026)
    p = tmp_path / "sub" / "config"
    cfg.save(p)
    assert DiaConfig.load(str(p) + ".json") == cfg
    assert DiaConfig.load(tmp_path / "missing.json") is None
    with pytest.raises(Exception):
        DiaConfig.model_validate({"model": {}, "data": {}})
    extra = cfg.model_dump()
    extra["training"] = {"lr": 1}                                   # the public HF config carries extra keys
    assert DiaConfig.model_validate(extra) == cfg


# ---- state ---------------------------------------------------------------------------------------------------
def test_kvcache_semantics():
    c = KVCache(4, 16, 128, torch.float32, torch.device("cpu"))
    assert c.k.shape == (2, 4, 16, 128) and c.current_idx == 0
    k1 = torch.randn(2, 4, 1, 128)
    ak, av = c.update(k1, k1 * 2)
    assert c.current_idx == 1 and ak.shape == (2, 4, 1, 128) and torch.equal(c.v[:, :, 0], k1[:, :, 0] * 2)
    kp = torch.randn(2, 4, 5, 128)
    rk, rv = c.prefill(kp, kp)
    assert rk is kp and c.current_idx == 4                          # n-1: the first decode step clobbers slot 4 (Q2)
    c.update(k1, k1)
    assert torch.equal(c.k[:, :, 4], k1[:, :, 0]) and c.current_idx == 5
    w = KVCache.from_kv(kp, kp)
    assert w.k is kp and w.current_idx == 0


def test_decoder_output_semantics():
    cfg = tiny_config()
    o = DecoderOutput.new(cfg, torch.device("cpu"))
    assert o.generated_tokens.shape == (cfg.data.audio_length, 9) and (o.generated_tokens == -1).all()
    o.prefill(torch.full((3, 9), 7, dtype=torch.int32), 1)
    assert o.get_tokens_at(0).shape == (9,) and o.get_tokens_at(0, 2).shape == (2, 9)
    o.generated_tokens[5, :4] = 1025
    o.update_one(torch.arange(9), 5, apply_mask=True)
    assert o.generated_tokens[5].tolist() == [1025] * 4 + [4, 5, 6, 7, 8]
    o.update_one(torch.arange(9), 5, apply_mask=False)
    assert o.generated_tokens[5].tolist() == list(range(9))


def test_masks_and_states_match_oracle():
    cfg = tiny_config()
    ids = O.encode_text(cfg, "[S1] hey")
    enc_in = torch.cat([torch.zeros_like(ids), ids])
    es = EncoderInferenceState.new(cfg, enc_in)
    pad = enc_in != 0
    assert torch.equal(es.attn_mask, O.create_attn_mask(pad, pad))
    assert es.positions.dtype == torch.float32 and es.positions.shape == (2, cfg.data.text_length)
    q = torch.rand(2, 6) > 0.5
    assert torch.equal(create_attn_mask(q, q, torch.device("cpu"), True), O.create_attn_mask(q, q, True))
    ds = DecoderInferenceState.new(cfg, es, torch.zeros(2, cfg.data.text_length, 256), [], torch.bfloat16)
    assert ds.dec_cross_attn_mask.shape == (2, 1, 1, cfg.data.text_length)
    assert not ds.dec_cross_attn_mask[0].any() and ds.dec_cross_attn_mask[1].sum() == 5
    assert all(c.k.dtype == torch.float32 for c in ds.self_attn_cache)               # KV is always fp32
    ds.prepare_step(7)
    assert ds.dec_positions.tolist() == [[7], [7]] and ds.dec_positions.dtype == torch.int32 and ds.step_from == 7
    ds.prepare_step(0, 4)
    assert ds.dec_positions.shape == (2, 4)


def test_module_tree_matches_reference_state_dict():
    from dia_tts_prune_b200.layers import DiaModel
    cfg = tiny_config()
    m = DiaModel(cfg, torch.float32)
    names = [n for n, _ in m.named_parameters()]
    assert names == O.param_names(cfg)
    shapes = O.param_shapes(cfg)
    assert all(tuple(p.shape) == shapes[n] for n, p in m.named_parameters())
    assert "decoder.layers.0.self_attention.rotary_emb.inv_freq" not in m.state_dict()   # non-persistent buffer
    with torch.device("meta"):
        big = DiaModel(dia_1_6b_config(), torch.bfloat16)
    sd = big.state_dict()
    assert len(sd) == 343
    assert sd["decoder.layers.3.mlp.wi_fused.weight"].dtype == torch.bfloat16
    assert sd["decoder.embeddings.0.weight"].dtype == torch.float32


def test_library_path_modules_match_oracle_on_cpu():
    """Encoder / cross-KV precompute / prefill (the once-per-utterance library path) vs the oracle."""
    from conftest import build_dia
    cfg = tiny_config()
    dia, sd = build_dia(cfg, 7)
    ids = O.encode_text(cfg, "[S1] Hello there. [S2] Hi. [S1]")
    enc_in = torch.cat([torch.zeros_like(ids), ids])
    es = EncoderInferenceState.new(cfg, enc_in)
    with torch.inference_mode():
        enc = dia.model.encoder(enc_in, es)
        pad = enc_in != 0
        ref = O.encoder_forward(sd, cfg, enc_in, es.positions, O.create_attn_mask(pad, pad))
        assert (enc - ref).abs().max() < 1e-5
        cross = dia.model.decoder.precompute_cross_attn_cache(enc, es.positions)
        rc = O.precompute_cross_kv(sd, cfg, ref, es.positions)
        assert cross[0].k.is_contiguous() and (cross[1].k - rc[1].k).abs().max() < 1e-5
        assert (cross[1].v - rc[1].v).abs().max() < 1e-5


def test_pruning_utils_masks():
    from dia_tts_prune_b200 import pruning_utils as PU
    from dia_tts_prune_b200.layers import DiaModel
    m = DiaModel(tiny_config(), torch.float32)
    SY.init_synthetic_(m.named_parameters(), 3)
    assert len(PU.get_prunable_modules(m)) == 2 * 6 + 2 * 10 + 1     # every DenseGeneral
    PU.apply_2to4_pruning(m)
    PU.make_pruning_permanent(m)
    w = m.decoder.layers[0].mlp.wo.weight
    assert ((w.reshape(-1, 4, w.shape[-1]) != 0).sum(1) <= 2).all()
    assert abs(PU.check_pruning_sparsity(m) - 0.5) < 0.01
    m2 = DiaModel(tiny_config(), torch.float32)
    SY.init_synthetic_(m2.named_parameters(), 3)
    PU.apply_structured_pruning(m2, 0.25, dim=0, n=2)
    PU.make_pruning_permanent(m2)
    wo = m2.decoder.layers[0].mlp.wo.weight
    assert int((wo.abs().sum(1) == 0).sum()) == wo.shape[0] // 4     # a quarter of the hidden neurons zeroed


def test_mlp_compaction_plan_drops_only_dead_neurons():
    """offline_prune.py --prune-dim 0 on mlp.wo zeroes hidden-neuron rows: the engine's weight stream keeps the live
    ones (padded to a width the kernel accepts) and the MLP output is unchanged."""
    from dia_tts_prune_b200 import pruning_utils as PU
    from dia_tts_prune_b200.layers import MlpBlock
    assert [PU.engine_mlp_width(n) for n in (1, 512, 513, 2048, 2049, 4096, 6000, 8192)] == \
        [512, 512, 1024, 2048, 4096, 4096, 6144, 8192]
    torch.manual_seed(3)
    D, F, L = 64, 4096, 2
    mlps = [MlpBlock(D, F, torch.float32) for _ in range(L)]
    params = {}
    for i, m in enumerate(mlps):
        with torch.no_grad():
            m.wi_fused.weight.normal_(0, D ** -0.5)
            m.wo.weight.normal_(0, F ** -0.5)
        params[f"layers.{i}.mlp.wi_fused.weight"] = m.wi_fused.weight
        params[f"layers.{i}.mlp.wo.weight"] = m.wo.weight
    assert PU.plan_mlp_compaction(params, L, F) is None                      # dense: nothing to drop
    import torch.nn.utils.prune as prune
    prune.ln_structured(mlps[0].wo, "weight", amount=0.6, n=2, dim=0); prune.remove(mlps[0].wo, "weight")
    prune.ln_structured(mlps[1].wo, "weight", amount=0.7, n=2, dim=0); prune.remove(mlps[1].wo, "weight")
    with torch.no_grad():
        mlps[1].wi_fused.weight[:, 0, 5] = 0.0                               # dead through its gate column
        mlps[1].wi_fused.weight[:, 1, 9] = 0.0                               # dead through its up column
    params = {k: v.detach() for k, v in params.items()}
    live1 = PU.mlp_live_neurons(params["layers.1.mlp.wi_fused.weight"], params["layers.1.mlp.wo.weight"])
    assert not live1[5] and not live1[9]
    width, keep = PU.plan_mlp_compaction(params, L, F)
    assert width == 2048                                                     # 40 % of 4096 = 1639 live -> 2048
    for i, idx in enumerate(keep):
        live = PU.mlp_live_neurons(params[f"layers.{i}.mlp.wi_fused.weight"], params[f"layers.{i}.mlp.wo.weight"])
        assert idx.numel() == width and torch.equal(idx, idx.unique()) and bool(live[idx].sum() == live.sum())
    small = PU.compact_mlp(params, (width, keep))
    x = torch.randn(2, 3, D)
    for i, m in enumerate(mlps):
        wi, wo = small[f"layers.{i}.mlp.wi_fused.weight"], small[f"layers.{i}.mlp.wo.weight"]
        assert wi.shape == (D, 2, width) and wo.shape == (width, D)
        gu = torch.tensordot(x, wi, dims=1)
        y = torch.tensordot(torch.nn.functional.silu(gu[..., 0, :]) * gu[..., 1, :], wo, dims=1)
        assert (y - m(x)).abs().max() < 1e-5


def test_2to4_detection_matches_the_mask_producer():
    """Decoder.engine() streams a checkpoint compressed only when EVERY dense kernel is 2:4 along its input axes."""
    from dia_tts_prune_b200 import pruning_utils as PU
    from dia_tts_prune_b200.layers import _contract_dims
    from dia_tts_prune_b200.model import Dia
    cfg = tiny_config()
    dia = Dia(cfg, "float32", torch.device("cpu"))
    SY.init_synthetic_(dia.model.named_parameters(), 3)
    dense = {n: p.detach() for n, p in dia.model.decoder.named_parameters() if SY.is_dense_kernel(n)}
    assert len(dense) == 10 * cfg.model.decoder.n_layer + 1     # incl. the cross-attention k/v projections (prepare only)
    assert not any(PU.is_2to4(t, _contract_dims(n)) for n, t in dense.items())
    PU.apply_2to4_pruning(dia.model.decoder)
    PU.make_pruning_permanent(dia.model)
    pruned = {n: p.detach() for n, p in dia.model.decoder.named_parameters() if SY.is_dense_kernel(n)}
    assert all(PU.is_2to4(t, _contract_dims(n)) for n, t in pruned.items())
    o = pruned["layers.0.self_attention.o_proj.weight"]                     # [heads, head_dim, D]: K = heads * head_dim
    assert _contract_dims("layers.0.self_attention.o_proj.weight") == 2 and PU.is_2to4(o, 2)
    assert abs(PU.check_pruning_sparsity(dia.model.decoder) - 0.5) < 0.01
    w = pruned["layers.1.mlp.wo.weight"].clone()
    w[0:4, 0] = 1.0                                                         # one group of one column with 4 non-zeros
    assert not PU.is_2to4(w, 1)


def test_synthetic_transcripts_are_deterministic():
    a = [SY.synthetic_transcript(i) for i in range(64)]
    assert a == [SY.synthetic_transcript(i) for i in range(64)]
    assert all(60 <= len(t.encode()) <= 200 and t.startswith("[S1]") for t in a)
    assert len(set(a)) > 50


def test_run_sharded_covers_every_utterance_once():
    texts = [SY.synthetic_transcript(i) for i in range(7)]
    seen = {}
    tot_frames = 0.0
    for rank in range(3):
        out, frames, secs = replicas.run_sharded(lambda t: torch.zeros((len(t), 9), dtype=torch.int32), texts, 3, rank)
        assert secs >= 0.0 and frames == sum(len(texts[i]) for i in out)
        assert not (set(out) & set(seen))
        seen.update(out)
        tot_frames += frames
    assert sorted(seen) == list(range(7)) and tot_frames == sum(len(t) for t in texts)
    assert all(seen[i].shape[0] == len(texts[i]) for i in range(7))


# ---- replicas (multi-GPU host logic) under gloo, world size 2 ----------------------------------------------------
def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = replicas.assign(7, world, rank)
    frames, secs = replicas.reduce_throughput(100.0 * len(mine), 1.0 + rank)
    allc = replicas.gather_codes([(i, torch.full((2, 9), i)) for i in mine])
    q.put((rank, mine, frames, secs, sorted(i for part in allc for i, _ in part)))
    dist.destroy_process_group()


def test_replicas_world_size_2_gloo():
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert res[0][1] == [0, 2, 4, 6] and res[1][1] == [1, 3, 5]
    for r in res:
        assert r[2] == 700.0 and r[3] == 2.0                        # frames summed, time = slowest rank
        assert r[4] == list(range(7))
    assert replicas.assign(3, 1, 0) == [0, 1, 2]
    assert replicas.reduce_throughput(5, 2) == (5.0, 2.0)           # identity when not distributed
    with pytest.raises(ValueError):
        replicas.assign(3, 2, 2)


def test_dia_accepts_every_reference_compute_dtype():
    """The reference's entry points default to float16 (cli.py --compute-dtype, app.py); every ComputeDtype value must
    construct (on the CPU the reference itself overrides to float32, dia/model.py:118-121)."""
    from dia_tts_prune_b200.config import tiny_config
    from dia_tts_prune_b200.model import ComputeDtype, Dia
    for dt in ComputeDtype:
        d = Dia(tiny_config(), dt.value, torch.device("cpu"))
        assert d.compute_dtype == torch.float32 and d.model is not None
    for dt in ("float16", "bfloat16", "float32"):
        from dia_tts_prune_b200.layers import DiaModel
        m = DiaModel(tiny_config(), {"float16": torch.float16, "bfloat16": torch.bfloat16, "float32": torch.float32}[dt])
        assert m.decoder.logits_dense.weight.dtype == {"float16": torch.float16, "bfloat16": torch.bfloat16,
                                                       "float32": torch.float32}[dt]


def test_canonicalize_rounds_once_and_widens_float16():
    import warnings
    from dia_tts_prune_b200 import layers as LY
    m = LY.DenseGeneral((64,), (128,), weight_dtype=torch.float16)
    with torch.no_grad():
        m.weight.copy_(torch.randn(64, 128) * 0.1)
    LY._ROUNDING_WARNED = False
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        assert LY.canonicalize_dense_kernel_(m) is True
    assert m.weight.dtype == torch.float32 and len(w) == 1
    assert torch.equal(m.weight, m.weight.to(torch.bfloat16).to(torch.float32))
    assert LY.canonicalize_dense_kernel_(m) is False                      # idempotent


def test_row_compaction_plan_is_exact():
    """K-row compaction of a stock `--prune-dim 0` checkpoint: the kept-row lists cover every live row, the maps are
    partial permutations, and contracting the compacted kernel with the gathered input equals the full product."""
    from dia_tts_prune_b200 import pruning_utils as PU
    from dia_tts_prune_b200.engine import decoder_tensor_names
    from dia_tts_prune_b200.model import Dia
    cfg = tiny_config(width=2)
    dia = Dia(cfg, "float32", torch.device("cpu"))
    SY.init_synthetic_(dia.model.named_parameters(), 3)
    PU.apply_structured_pruning(dia.model, 0.5, dim=0, n=2)
    PU.make_pruning_permanent(dia.model)
    sd = dict(dia.model.decoder.named_parameters())
    tensors = {n: sd[n].detach() for n in decoder_tensor_names(cfg)}
    L = cfg.model.decoder.n_layer
    plan = PU.plan_row_compaction(tensors, L)
    # cross-q, mlp-in and the logits head lose half of their 1024 input rows, the attention output projections half of
    # their heads; the fused q/k/v projection keeps a row that is live in any of the three (87.5 %): not worth it
    assert set(plan) == {1, 2, 3, 4, 6} and all(w == 512 for w, _ in plan.values())
    compact, maps = PU.compact_rows(tensors, plan)
    g = torch.Generator().manual_seed(0)
    for fam, (width, keep) in plan.items():
        m = maps[fam]
        assert m.dtype == torch.int32 and m.shape == ((1 if fam == 6 else L), 1024)
        for l in range(m.shape[0]):
            pos = m[l][m[l] >= 0]
            assert sorted(pos.tolist()) == list(range(width))
        name = "logits_dense.weight" if fam == 6 else f"layers.1.{PU.K_FAMILIES[fam][0]}"
        w_full, w_c = tensors[name].reshape(1024, -1), compact[name].reshape(width, -1)
        x = torch.randn(1024, generator=g)
        row = m[0 if fam == 6 else 1]
        xc = torch.zeros(width)
        xc[row[row >= 0].long()] = x[row >= 0]
        assert torch.allclose(xc @ w_c, x @ w_full, atol=1e-5)
    # an unpruned model: nothing to compact
    dia2 = Dia(cfg, "float32", torch.device("cpu"))
    SY.init_synthetic_(dia2.model.named_parameters(), 3)
    sd2 = dict(dia2.model.decoder.named_parameters())
    assert PU.plan_row_compaction({n: sd2[n].detach() for n in decoder_tensor_names(cfg)}, L) == {}
