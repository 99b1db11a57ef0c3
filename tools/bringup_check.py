#!/usr/bin/env python
"""Bring-up diagnostics on a GPU box: stage-by-stage diff of the step kernel against the CPU oracle.

Test infrastructure (imports oracle/).  Usage: python tools/bringup_check.py [--full] [--steps N]
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from dia_tts_prune_b200 import _lib                                    # noqa: E402
from dia_tts_prune_b200.config import dia_1_6b_config, tiny_config     # noqa: E402
from dia_tts_prune_b200.model import Dia                               # noqa: E402
from dia_tts_prune_b200 import synthetic as SY                         # noqa: E402
from oracle import dia_oracle as O                                     # noqa: E402


def err(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return (a - b).abs().max().item(), b.abs().max().item()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--full", action="store_true")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--seed", type=int, default=7)
    a = ap.parse_args()
    cfg = dia_1_6b_config() if a.full else tiny_config()
    dev = torch.device("cuda:0")
    t0 = time.time()
    dia = Dia(cfg, "float32", torch.device("cpu"))
    SY.init_synthetic_(dia.model.named_parameters(), a.seed)
    sd = {k: v.detach().clone() for k, v in dia.model.named_parameters()}
    print(f"weights ready {time.time()-t0:.1f}s fingerprint {SY.weights_fingerprint(sd, [n for n in sd if 'layers.0.' in n])[:16]}")
    dia.device = dev
    dia.live_text_only = False           # compare the full encoder / cross-KV tensors with the oracle
    dia.model.to(dev)
    dia.model.eval()
    text = "[S1] Hello there. [S2] Hi."
    eff = O.effective_text(text, None)
    with torch.inference_mode():
        dec_state, dec_out = dia._prepare_generation(eff, None, False)
        st_o, grid_o, P0 = O.prepare_generation(sd, cfg, eff, None, dead_cross_kv=False)
        print("enc_out err", err(dec_state.enc_out, st_o.enc_out))
        print("cross k L0 err", err(dec_state.cross_attn_cache[0].k, st_o.cross_cache[0].k), "text_len", dec_state.text_len)
        eng = dia.model.decoder._engine_for(dec_state)
        print("engine ctas", eng.n_ctas, "weight stream MB", eng.weight_stream_bytes / 1e6)
        L = cfg.model.decoder.n_layer
        grid = grid_o.clone()
        for step in range(a.steps):
            cur = P0 + step
            slot = step
            toks = grid[cur - 1].unsqueeze(0).expand(2, -1).contiguous()
            # ---- oracle, layer by layer
            st_o.prepare_step(cur)
            x = O.embed_sum(sd, cfg, toks.unsqueeze(1))
            xs_o = [x.clone()]
            for i in range(L):
                x = O.decoder_layer(sd, cfg, i, x, st_o, prefill=False, dead_cross_kv=False)
                xs_o.append(x.clone())
            xn = O.rms_norm(x, sd["decoder.norm.weight"], cfg.model.normalization_layer_epsilon)
            logits_o = O.dense(xn, sd["decoder.logits_dense.weight"])[:, -1]
            # ---- kernel, stage by stage (idempotent: same slot every time)
            tk = toks.to(dev)
            eng.run_stages(tk, 0, 1, cur, slot)
            torch.cuda.synchronize()
            print(f"step {step} embed err", err(eng.read_buffer(_lib.BUF_X), xs_o[0][:, 0]))
            if step == 0 or not a.full:
                for i in range(min(L, 2)):
                    eng.run_stages(tk, 0, 1 + 8 * (i + 1), cur, slot)
                    torch.cuda.synchronize()
                    print(f"  layer {i} out err", err(eng.read_buffer(_lib.BUF_X), xs_o[i + 1][:, 0]))
            logits = eng.decode_step(tk, cur, slot)
            torch.cuda.synchronize()
            print(f"step {step} x_final err", err(eng.read_buffer(_lib.BUF_X), xs_o[L][:, 0]),
                  " logits err", err(logits, logits_o))
            for c in dec_state.self_attn_cache:
                c.current_idx = slot + 1
            print("   self K slot err L0", err(dec_state.self_attn_cache[0].k[:, :, slot], st_o.self_cache[0].k[:, :, slot]),
                  "L-1", err(dec_state.self_attn_cache[L - 1].k[:, :, slot], st_o.self_cache[L - 1].k[:, :, slot]))
            guided = O.cfg_combine_and_mask(cfg, logits_o.clone(), 3.0)
            pred_o = torch.argmax(guided, -1)
            pred = eng.head_sample(logits, 3.0, 0.0, 0.95, 35)
            print("   pred match", bool((pred.cpu() == pred_o).all()), pred.cpu().tolist(), pred_o.tolist())
            grid[cur] = torch.where(grid[cur] == -1, pred_o.int(), grid[cur]) if step < 14 else pred_o.int()
        # ---- fused generate vs oracle generate
        t0 = time.time()
        tr = O.generate(sd, cfg, text, max_tokens=40 if not a.full else 20, temperature=0.0, dead_cross_kv=False)
        print(f"oracle generate {time.time()-t0:.1f}s")
        t0 = time.time()
        out = dia.generate(text, max_tokens=40 if not a.full else 20, temperature=0.0, output="codes")
        torch.cuda.synchronize()
        print(f"kernel generate {time.time()-t0:.3f}s stats {dia.last_stats}")
        same = torch.equal(dia.last_codes.cpu(), tr.codes)
        print("greedy codes equal:", same, "min margin", torch.stack(tr.margins).min().item())
        if not same:
            d = (dia.last_codes.cpu() != tr.codes).nonzero()
            print("first diffs", d[:5].tolist())
        print("finalized equal:", torch.equal(out.cpu(), O.finalize_codes(cfg, tr.codes).to(torch.int32)))


if __name__ == "__main__":
    main()
