#!/usr/bin/env python
"""Repeat a greedy 8-utterance generate_batch and compare the code streams of every repetition with the first one: the
batched kernel is deterministic by construction, any difference is a race.  python tools/batch_determinism.py [--reps 8]"""
import argparse, os, sys
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from dia_tts_prune_b200 import synthetic as SY
from dia_tts_prune_b200.config import dia_1_6b_config
from dia_tts_prune_b200.model import Dia

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=8)
ap.add_argument("--tokens", type=int, default=300)
ap.add_argument("--utts", type=int, default=8)
a = ap.parse_args()
cfg = dia_1_6b_config()
dia = Dia(cfg, "float32", torch.device("cpu"))
SY.init_synthetic_(dia.model.named_parameters(), 5)
SY.cast_dense_kernels_(dia.model, torch.bfloat16)
dia.device = torch.device("cuda:0")
dia.model.to(dia.device).eval()
dia.batch_min_utterances = 1
texts = [SY.synthetic_transcript(i) for i in range(a.utts)]
first = None
bad = 0
for rep in range(a.reps):
    try:
        dia.generate_batch(texts, max_tokens=a.tokens, temperature=0.0, max_utterances=a.utts)
    except Exception as ex:
        from collections import Counter
        head, where = dia.model.decoder._bengine.last_device_error(full=True)
        print(f"rep {rep}: {ex.__class__.__name__}: device error words {head}")
        print("   sites:", Counter(where.values()).most_common(10))
        byw = {}
        for (b, w), v in where.items():
            byw.setdefault((w, v), []).append(b)
        for (w, v), bl in sorted(byw.items()):
            print(f"    warp {w} site {v}: {len(bl)} blocks {bl[:6]}..{bl[-3:]}")
        sys.exit(1)
    cur = [c.cpu().clone() for c in dia.last_batch_codes]
    if first is None:
        first = cur
        continue
    for u, (x, y) in enumerate(zip(first, cur)):
        if not torch.equal(x, y):
            d = (x != y).nonzero()
            bad += 1
            print(f"rep {rep} utterance {u}: first difference at row {d[0, 0].item()} channel {d[0, 1].item()} ({d.shape[0]} cells differ)")
print(f"{a.reps} repetitions, {bad} differing (repetition, utterance) pairs")
