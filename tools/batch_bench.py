#!/usr/bin/env python
"""Step time of the batched (N utterances per launch) kernel at a given context depth.

    python tools/batch_bench.py [--utts 1 2 4 8] [--slot 1500] [--steps 64] [--reps 5] [--tiny]
Prints us/step and frames/s per GPU for each N (frames/step = N)."""
import argparse, os, sys, time
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from dia_tts_prune_b200 import synthetic as SY
from dia_tts_prune_b200.config import dia_1_6b_config, tiny_config
from dia_tts_prune_b200.model import Dia

ap = argparse.ArgumentParser()
ap.add_argument("--utts", type=int, nargs="+", default=[1, 2, 4, 8])
ap.add_argument("--slot", type=int, default=1500)
ap.add_argument("--steps", type=int, default=64)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--tiny", action="store_true")
ap.add_argument("--profile", action="store_true", help="print where CTA 0's MMA thread and math thread 0 spend their clocks")
a = ap.parse_args()
cfg = tiny_config() if a.tiny else dia_1_6b_config()
dev = torch.device("cuda:0")
dia = Dia(cfg, "float32", torch.device("cpu"))
SY.init_synthetic_(dia.model.named_parameters(), 5)
with torch.no_grad():
    dia.model.decoder.logits_dense.weight[:, 0, 1024] = 0.0
SY.cast_dense_kernels_(dia.model, torch.bfloat16)
dia.device = dev
dia.model.to(dev).eval()
slot = min(a.slot, cfg.data.audio_length - a.steps - 2)
with torch.inference_mode():
    eng = dia.model.decoder.batch_engine(max(a.utts))
    prepared = [dia._prepare_generation(dia._effective_text(SY.DEFAULT_TRANSCRIPT, None), None, False) for _ in range(max(a.utts))]
    for u, (st, out) in enumerate(prepared):
        eng.bind(u, st.self_attn_cache, st.cross_attn_cache, st.text_len)
        out.generated_tokens[: slot + a.steps + 2] = 7
    for U in a.utts:
        best = None
        for rep in range(a.reps):
            eng.generate_begin([o.generated_tokens for _, o in prepared[:U]], [slot + 1] * U, [slot] * U, cfg.data.audio_length,
                               3.0, 1.3, 0.95, 35, list(range(U)))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            eng.generate_steps(a.steps)
            e1.record()
            try:
                torch.cuda.synchronize()
            except Exception as ex:
                print(f"N={U}: launch failed: {ex.__class__.__name__}; device error words {eng.last_device_error()}")
                sys.exit(1)
            st = eng.status()
            assert all(s.steps_run == a.steps for s in st), [s.steps_run for s in st]
            us = e0.elapsed_time(e1) * 1000 / a.steps
            best = us if best is None else min(best, us)
        if a.profile:
            eng.enable_timing(True)
            eng.generate_begin([o.generated_tokens for _, o in prepared[:U]], [slot + 1] * U, [slot] * U, cfg.data.audio_length,
                               3.0, 1.3, 0.95, 35, list(range(U)))
            eng.generate_steps(a.steps)
            torch.cuda.synchronize()
            pr = eng.read_profile()
            eng.enable_timing(False)
            print("   profile (us/step): " + "  ".join(f"{k}={v / 1965.0 / a.steps:.1f}" for k, v in pr.items()))
        print(f"N={U}: {best:8.1f} us/step  {U / best * 1e6:9.1f} frames/s per GPU  (slot {slot}, {a.steps} steps/launch, best of {a.reps})")
