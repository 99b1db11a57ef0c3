import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import build_dia
from dia_tts_prune_b200 import pruning_utils as PU
from dia_tts_prune_b200.config import tiny_config
from oracle import dia_oracle as O
cfg = tiny_config(width=2)
dia, _ = build_dia(cfg, 11)
PU.apply_structured_pruning(dia.model, 0.5, dim=0, n=2)
PU.make_pruning_permanent(dia.model)
sd = {k: v.detach().clone() for k, v in dia.model.named_parameters()}
dia.device = torch.device("cuda:0"); dia.model.to(dia.device)
text = "[S1] Rows dropped. [S2] Same logits."
for compact in (True, False):
    dia.model.decoder.compact_pruned_rows = compact
    dia.model.decoder.invalidate_engine()
    with torch.inference_mode():
        st2, _ = dia._prepare_generation(dia._effective_text(text, None), None, False)
    st_o, _, _ = O.prepare_generation(sd, cfg, O.effective_text(text, None), None, dead_cross_kv=False)
    st2.prepare_step(1); st_o.prepare_step(1)
    x = torch.randn(2, 1, cfg.model.decoder.n_embd, generator=torch.Generator().manual_seed(2))
    xg, xo = x.cuda(), x.clone()
    with torch.inference_mode():
        for i, layer in enumerate(dia.model.decoder.layers):
            xg = layer(xg, st2, self_attn_cache=st2.self_attn_cache[i], cross_attn_cache=st2.cross_attn_cache[i])
            xo = O.decoder_layer(sd, cfg, i, xo, st_o, prefill=False, dead_cross_kv=False)
            d = (xg.cpu() - xo).abs()
            print(f"compact={compact} layer {i}: row0 err {d[0].max():.3e} row1 err {d[1].max():.3e}; k_rows {dia.model.decoder.engine().k_rows}")
