#!/usr/bin/env python
"""Per-stage time breakdown of the persistent step kernel (in-kernel SM-clock stamps of CTA 0).

    python tools/stage_profile.py [--slot 0] [--steps 8] [--tiny]
Prints, per stage kind, the mean work time and the mean barrier wait of CTA 0, in microseconds.
"""
import argparse
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from dia_tts_prune_b200 import synthetic as SY                              # noqa: E402
from dia_tts_prune_b200.config import dia_1_6b_config, tiny_config          # noqa: E402
from dia_tts_prune_b200.model import Dia                                    # noqa: E402

NAMES = ["qkv", "sattn", "so", "cq", "cattn", "co", "wi", "wo"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--slot", type=int, default=0)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--tiny", action="store_true")
    ap.add_argument("--greedy", action="store_true")
    ap.add_argument("--cta", type=int, default=0, help="the CTA whose stamps are shown")
    ap.add_argument("--ctas", action="store_true", help="also print the per-CTA spread of stage end times")
    ap.add_argument("--mlp-prune", type=float, default=0.0, help="structurally prune mlp.wo (dim 0) by this amount first")
    a = ap.parse_args()
    cfg = tiny_config() if a.tiny else dia_1_6b_config()
    dev = torch.device("cuda:0")
    dia = Dia(cfg, "float32", torch.device("cpu"))
    SY.init_synthetic_(dia.model.named_parameters(), 5)
    with torch.no_grad():
        dia.model.decoder.logits_dense.weight[:, 0, 1024] = 0.0
    if a.mlp_prune > 0:
        import torch.nn.utils.prune as prune
        from dia_tts_prune_b200 import pruning_utils as PU
        for layer in dia.model.decoder.layers:
            prune.ln_structured(layer.mlp.wo, "weight", amount=a.mlp_prune, n=2, dim=0)
        PU.make_pruning_permanent(dia.model)
    SY.cast_dense_kernels_(dia.model, torch.bfloat16)
    dia.device = dev
    dia.model.to(dev).eval()
    with torch.inference_mode():
        st, out = dia._prepare_generation(dia._effective_text(SY.DEFAULT_TRANSCRIPT, None), None, False)
        eng = dia.model.decoder._engine_for(st)
        mhz = torch.cuda.clock_rate() if hasattr(torch.cuda, "clock_rate") else 1965
        L = cfg.model.decoder.n_layer
        S = 8 * L + 3
        # fill the grid so that any starting slot has a valid input row
        out.generated_tokens[: a.slot + a.steps + 2] = 7
        for it in range(2):
            eng.enable_timing(it == 1, a.cta)
            eng.generate_begin(out.generated_tokens, a.slot + 1, a.slot, cfg.data.audio_length, 3.0,
                               0.0 if a.greedy else 1.3, 0.95, 35, 1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            eng.generate_steps(a.steps)
            e1.record()
            try:
                torch.cuda.synchronize()
            except Exception as ex:
                print("launch failed:", ex.__class__.__name__, "device error words", eng.last_device_error())
                raise
        t = eng.read_timing(a.steps).double()
        print(f"launch of {a.steps} steps from slot {a.slot}: {e0.elapsed_time(e1) * 1000 / a.steps:.1f} us/step (event)")
        t = t[1:]                                                      # skip the cold first step
        us = 1.0 / mhz
        # no barriers: a stage's cost on CTA 0 = start of the next stage - its own start
        starts = t[:, :, 0].reshape(-1)
        dur = torch.cat([starts[1:] - starts[:-1], (t[-1, -1, 4] - t[-1, -1, 0]).reshape(1)]).reshape(t.shape[0], S)
        tot = dur.sum(1).mean().item() * us
        print(f"clock {mhz} MHz; per-step total from stamps {tot:.1f} us (steps 1..{a.steps - 1})")
        print(f"{'stage':8s} {'setup':>7s} {'wait in':>8s} {'loop':>7s} {'reduce':>7s} {'epi':>7s} {'stage':>7s} {'x/step':>6s} {'total':>8s}")
        rows = [("embed", [0])] + [(NAMES[j], [1 + 8 * l + j for l in range(L)]) for j in range(8)] + \
               [("logits", [8 * L + 1]), ("sample", [8 * L + 2])]
        for name, idx in rows:
            x = t[:, idx, :]
            d = lambda i, j: ((x[..., i] - x[..., j]).mean().item() * us)     # noqa: E731
            stage = dur[:, idx].mean().item() * us
            if name not in ("embed", "sattn", "cattn", "sample"):
                print(f"{name:8s} {d(1, 0):7.2f} {d(6, 1):8.2f} {d(2, 6):7.2f} {d(3, 2):7.2f} {d(4, 3):7.2f} {stage:7.2f} "
                      f"{len(idx):6d} {stage * len(idx):8.1f}   epi: sum {d(5, 3):5.2f}  warp7: loop end {d(8, 0):5.2f} at barrier {d(10, 0):5.2f} | warp0 at barrier {d(3, 0):5.2f}")
            elif name in ("sattn", "cattn"):
                print(f"{name:8s} {'':7s} {d(1, 0):8.2f} {d(2, 1):7.2f} {'':7s} {d(4, 2):7.2f} {stage:7.2f} "
                      f"{len(idx):6d} {stage * len(idx):8.1f}   warp0 tile0: K wait {d(11, 1):5.2f} scores {d(12, 11):5.2f} "
                      f"softmax+V wait {d(13, 12):5.2f} PV {d(14, 13):5.2f} to loop end {d(2, 14):5.2f} | merge+store {d(15, 2):5.2f}"
                      + (f" partials in {d(5, 15):5.2f} barrier {d(6, 5):5.2f} combine+store {d(7, 6):5.2f} fence+barrier {d(4, 7):5.2f}" if name == "sattn" else ""))
            elif name == "sample":
                print(f"{name:8s} {'':7s} {'':8s} {'':7s} {'':7s} {'':7s} {stage:7.2f} {len(idx):6d} {stage * len(idx):8.1f}   "
                      f"logits in {d(1, 0):5.2f} radix select {d(2, 1):5.2f} survivors {d(6, 2):5.2f} rank sort {d(7, 6):5.2f} "
                      f"softmax+top-p {d(8, 7):5.2f} draw+publish {d(3, 8):5.2f} all preds {d(5, 3):5.2f} state machine {d(4, 5):5.2f}")
            else:
                print(f"{name:8s} {'':7s} {'':8s} {'':7s} {'':7s} {'':7s} {stage:7.2f} {len(idx):6d} {stage * len(idx):8.1f}")
        if a.ctas:
            ct = eng.read_cta_timing().double()                          # [S, G] ns, step 1
            end = ct.max(1).values
            print("per-stage spread over CTAs (step 1): stage, last CTA finish - first CTA finish [us], slowest CTA")
            for probe in ("wi", "sattn", "wo"):
                j = NAMES.index(probe)
                for l in (3, 9):
                    sidx = 1 + 8 * l + j
                    rel = (ct[sidx] - end[sidx - 1]) / 1000.0          # finish time after the previous stage completed everywhere
                    order = torch.argsort(rel)
                    print(f"  {probe} layer {l}: finish after prev stage [us] min {rel.min():.2f} median {rel.median():.2f} max {rel.max():.2f};"
                          f" slowest CTAs {order[-6:].tolist()} fastest {order[:6].tolist()}")
            for name, idx in rows:
                sp = (ct[idx].max(1).values - ct[idx].min(1).values).mean().item() / 1000
                print(f"  {name:8s} spread {sp:6.2f}  stage-to-stage {((end[idx] - end[[max(i - 1, 0) for i in idx]]).mean().item()) / 1000:6.2f}")


if __name__ == "__main__":
    main()
