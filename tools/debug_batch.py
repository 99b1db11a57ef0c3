#!/usr/bin/env python
"""Bring-up of the batched kernel on a GPU box: one tiny / full decode_step per utterance count, device error words on failure,
logits against the oracle.  Test infrastructure (imports oracle/).  python tools/debug_batch.py [--full] [--utts 1 2 3]"""
import argparse, os, sys
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from dia_tts_prune_b200 import synthetic as SY
from dia_tts_prune_b200.config import dia_1_6b_config, tiny_config
from dia_tts_prune_b200.model import Dia
from oracle import dia_oracle as O

ap = argparse.ArgumentParser()
ap.add_argument("--full", action="store_true")
ap.add_argument("--utts", type=int, nargs="+", default=[1, 2, 3])
ap.add_argument("--steps", type=int, default=3)
a = ap.parse_args()
cfg = dia_1_6b_config() if a.full else tiny_config()
dev = torch.device("cuda:0")
dia = Dia(cfg, "float32", torch.device("cpu"))
SY.init_synthetic_(dia.model.named_parameters(), 7)
SY.cast_dense_kernels_(dia.model, torch.bfloat16)
sd = {k: v.detach().clone().float() for k, v in dia.model.named_parameters()}
dia.device = dev
dia.model.to(dev).eval()
TEXTS = ["[S1] Hello there. [S2] Hi.", "[S1] A second utterance, a little longer than the first. [S2] Yes.",
         "[S2] Third one starts with speaker two. [S1] Fine."]
with torch.inference_mode():
    eng = dia.model.decoder.batch_engine(4)
    for U in a.utts:
        states, ostates = [], []
        for u in range(U):
            st, out = dia._prepare_generation(dia._effective_text(TEXTS[u], None), None, False)
            eng.bind(u, st.self_attn_cache, st.cross_attn_cache, st.text_len)
            states.append(st)
            ostates.append(O.prepare_generation(sd, cfg, O.effective_text(TEXTS[u], None), None, dead_cross_kv=False)[0])
        g = torch.Generator().manual_seed(U)
        for cur in range(1, 1 + a.steps):
            toks = torch.randint(0, 1024, (U, 9), generator=g, dtype=torch.int32)
            try:
                lg = eng.decode_step(toks.cuda(), [cur] * U, [cur - 1] * U)
                torch.cuda.synchronize()
            except Exception as ex:
                head, where = eng.last_device_error(full=True)
                print(f"U={U} step {cur}: launch failed: {ex.__class__.__name__}; head {head}")
                from collections import Counter
                cnt = Counter(where.values())
                print("   sites:", cnt.most_common(12))
                common = {v for v, _ in cnt.most_common(2)}
                byw = {}
                for (b, w), v in where.items():
                    byw.setdefault((w, v), []).append(b)
                for (w, v), bl in sorted(byw.items()):
                    print(f"    warp {w} site {v}: {len(bl)} blocks {bl[:6]}..{bl[-3:]}")
                sys.exit(1)
            lg = lg.cpu()
            worst = 0.0
            for u in range(U):
                ostates[u].prepare_step(cur)
                lo = O.decoder_forward(sd, cfg, toks[u].long().unsqueeze(0).unsqueeze(0).expand(2, 1, -1), ostates[u],
                                       prefill=False, dead_cross_kv=False)[:, 0]
                worst = max(worst, (lg[2 * u: 2 * u + 2] - lo).abs().max().item())
            print(f"U={U} step {cur}: max-abs logits error {worst:.3e}")
