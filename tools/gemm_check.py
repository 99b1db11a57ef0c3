#!/usr/bin/env python
"""tcgen05 GEMM against torch (float64 reference), a few shapes incl. an M tail; prints max-abs error and TFLOP/s."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dia_tts_prune_b200 import engine as E

torch.manual_seed(0)
for (M, N, K) in [(128, 128, 64), (128, 128, 256), (200, 256, 512), (1722, 2048, 2048), (2048, 16384, 2048), (1722, 2048, 8192)]:
    x = torch.randn(M, K, device="cuda")
    w = (torch.randn(K, N, device="cuda") * K ** -0.5).to(torch.bfloat16)
    wt = E.dense_prepare_weight(w)
    assert torch.equal(wt, w.t().contiguous())
    y = E.dense_forward(x, wt)
    torch.cuda.synchronize()
    ref = (x.double() @ w.double()).float()
    err = (y - ref).abs().max().item()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        y = E.dense_forward(x, wt)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    t32 = time.time()
    print(f"M={M} N={N} K={K}: max-abs err {err:.2e} (|ref| max {ref.abs().max():.2f}), {ms*1e3:.0f} us/call incl. split, "
          f"{3 * 2 * M * N * K / ms / 1e9:.1f} TFLOP/s issued ({2 * M * N * K / ms / 1e9:.1f} effective)")
