#!/usr/bin/env python
"""bench.py - decoded audio frames/s of the Dia-1.6B decode path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): Dia-1.6B, bf16 weights, ONE transcript, CFG batch 2, a full
3072-token generation = 3071 autoregressive decode steps.  A bench "step" is one such generation.
With N > 1 GPUs every rank runs its own utterance (replicas, no collective on the data path:
SURVEY.md 8(e)); `value` is the sum over ranks divided by the slowest rank's time (weak scaling).

  value  frames/s of the decode loop with everything already resident in HBM (encoder output,
         cross-KV, token grid), timed on the device with CUDA events on the launching stream.
  e2e    frames/s through the public API, `Dia.generate(text, output="codes")` + `.cpu()`: text in host
         memory in, codes in host memory out; encoder, cross-KV precompute, H2D/D2H inside the region.
  roofline     the step kernel: algorithmic bytes of the steps in a launch / event-timed launch duration,
               against the measured HBM copy bandwidth in MEASURED_PEAKS.json.
  cpu_baseline the CPU port of the reference's own algorithm (oracle/, kind "port": the reference is
               Python and cannot travel to the GPU box) timed on this box's host cores.

`--impl reference` times that CPU port alone (rank 0 only).
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

METRIC = "decoded_audio_frames_per_s"
UNIT = "frames/s"
MAX_TOKENS = 3072
WEIGHT_SEED = 5
FRAME_RATE = 44100 / 512          # 86.13 frames per second of audio


VARIANTS = {
    "dense": "Dia-1.6B",
    "mlp-pruned": "Dia-1.6B with mlp.wo structurally pruned along dim 0 (offline_prune.py --prune-dim 0, L2 norm; "
                  "BASELINE.json configs[3] 'reduced MLP width')",
    "2to4": "Dia-1.6B with every dense kernel 2:4-pruned along K (BASELINE.json configs[3], unstructured-sparse variant)",
    "structured": "Dia-1.6B after the stock offline_prune.py call: apply_structured_pruning(model, amount, dim=0, n=2) on "
                  "EVERY DenseGeneral (input rows zeroed; the engine drops them: K-row compaction + reduced MLP width)",
}


def workload_config(text_len: int, world: int, variant: str = "dense", amount: float = 0.0) -> dict:
    """The `config` object both arms print (BASELINE.json configs[1], one utterance per GPU)."""
    model = VARIANTS[variant] + (f", amount {amount}" if variant in ("mlp-pruned", "structured") else "")
    return {"workload": f"{model} bf16 weights, single transcript per GPU, CFG batch 2, full 3072-token "
                        "generation (3071 decode steps), reference default sampling T=1.3 top_p=0.95 top_k=35",
            "precision": "bf16 weights; fp32 activations, accumulation, softmax and KV cache",
            "weights": f"random-init seed {WEIGHT_SEED}, channel-0 EOS column zeroed so no early EOS",
            "l2": "inputs larger than L2: 2.53 GB of weights streamed per decode step",
            "frame": "1 frame = 1 decode step = 9 codes = 512 samples @ 44.1 kHz",
            "text_len": text_len, "parallelism": f"replicas x{world}, no data-path collective"}


def algorithmic_bytes(cfg, slot: int, text_len: int, kv_elem: int = 4, n_hidden: int | None = None,
                      weight_scale: float = 1.0, dense_bytes: int | None = None) -> int:
    """bf16 weights of the 18 layers + logits head, fp32 norm weights, 9 embedding rows,
    plus self-KV rows read / appended and the valid cross-KV rows of the conditional row.  Pruned variants:
    ``n_hidden`` = live MLP width of a structurally pruned model; ``weight_scale`` = 0.5625 for 2:4 (half the
    values + 2 bits of metadata per kept value)."""
    d, e = cfg.model.decoder, cfg.model.encoder
    D, F, hd = d.n_embd, n_hidden or d.n_hidden, 128
    nq, nkv, nc = d.gqa_query_heads * hd, d.kv_heads * hd, d.cross_query_heads * hd
    per_layer = D * (nq + 2 * nkv) + nq * D + D * nc + nc * D + D * 2 * F + F * D
    w = int(2 * weight_scale * (d.n_layer * per_layer + D * cfg.data.channels * cfg.model.tgt_vocab_size))
    if dense_bytes is not None:        # row-compacted (structurally pruned) model: the live kernels as the engine streams them
        w = int(dense_bytes)
    w += 4 * (3 * d.n_layer + 1) * D + 4 * cfg.data.channels * D
    kv_row = d.n_layer * 2 * 2 * nkv * kv_elem            # K and V, both CFG rows, all layers, one slot
    cross = d.n_layer * 2 * nc * kv_elem * text_len       # K and V, conditional row only
    return w + kv_row * (slot + 1) + cross


class ClockSampler:
    """nvidia-smi sampled during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        load = [x for x in sm if mx and x > 0.5 * mx] or sm
        return {"sm_mhz": statistics.median(load) if load else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peak() -> tuple[float, str]:
    try:
        with open(os.path.join(REPO, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def build_cpu_model(seed: int, suppress_eos: bool):
    """fp32 CPU DiaModel with the shared synthetic recipe; returns (Dia, state dict clone or None)."""
    import torch
    from dia_tts_prune_b200.config import dia_1_6b_config
    from dia_tts_prune_b200.model import Dia
    from dia_tts_prune_b200 import synthetic as SY
    cfg = dia_1_6b_config()
    dia = Dia(cfg, "float32", torch.device("cpu"))
    SY.init_synthetic_(dia.model.named_parameters(), seed)
    if suppress_eos:
        # random weights would emit EOS on channel 0 within ~1000 steps and end the utterance early;
        # a zero EOS column keeps all 3071 steps running
        with torch.no_grad():
            dia.model.decoder.logits_dense.weight[:, 0, cfg.data.audio_eos_value] = 0.0
    return dia, cfg


def cpu_port_frames_per_s(sd, cfg, text: str, n_steps: int, threads: int, lean_steps: int = 0) -> dict:
    """The oracle port of the reference's decode loop, timed on the host (dead cross-attention K/V
    re-projection kept, as shipped: dia/layers.py:274-275).  ``lean_steps`` > 0 also times the same loop with that
    dead projection skipped (identical outputs; BASELINE.md section 3's "useful-work" CPU figure)."""
    import torch
    from oracle import dia_oracle as O
    torch.set_num_threads(threads)
    t0 = time.perf_counter()
    tr = O.generate(sd, cfg, text, max_tokens=1 + n_steps, temperature=0.0, cfg_scale=3.0, dead_cross_kv=True,
                    time_steps=True)
    total = time.perf_counter() - t0
    loop = sum(tr.step_seconds)
    out = {"value": len(tr.step_seconds) / loop, "unit": UNIT, "cores": threads, "kind": "port",
           "sample": f"{len(tr.step_seconds)} greedy decode steps (cache slots 0..{len(tr.step_seconds) - 1}) of the same "
                     f"Dia-1.6B fp32 workload on the host (CFG batch 2, dead cross-K/V projection kept as shipped), loop "
                     f"{loop:.1f}s; one-time encoder+cross-KV prepare {total - loop:.1f}s excluded"}
    if lean_steps > 0:
        tr = O.generate(sd, cfg, text, max_tokens=1 + lean_steps, temperature=0.0, cfg_scale=3.0, dead_cross_kv=False,
                        time_steps=True)
        out["lean_value"] = len(tr.step_seconds) / sum(tr.step_seconds)
        out["lean_note"] = (f"same loop over {len(tr.step_seconds)} steps with the reference's dead per-step cross-attention "
                            "K/V re-projection skipped (outputs identical)")
    return out


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from dia_tts_prune_b200 import synthetic as SY
    threads = os.cpu_count() or 1
    dia, cfg = build_cpu_model(WEIGHT_SEED, suppress_eos=True)
    sd = {k: v.detach() for k, v in dia.model.named_parameters()}
    eff = dia._effective_text(SY.DEFAULT_TRANSCRIPT, None)                 # speaker tags are one byte each
    text_len = len(eff.encode("utf-8").replace(b"[S1]", b"\x01").replace(b"[S2]", b"\x02"))
    per_step = max(2, args.ref_decode_steps)
    vals, t_all = [], time.perf_counter()
    for i in range(args.warmup + args.steps):
        r = cpu_port_frames_per_s(sd, cfg, SY.DEFAULT_TRANSCRIPT, per_step, threads)
        if i >= args.warmup:
            vals.append(r)
    v = statistics.mean(x["value"] for x in vals)
    line = {"metric": METRIC, "value": v, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000.0 * per_step / v, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(text_len, max(1, args.gpus)),
            "note": f"reference arm: the CPU port of the reference's own fp32 loop; every bench step is a bounded sample of "
                    f"{per_step} greedy decode steps of that workload (precision fp32 on the host, not bf16)",
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": vals[-1]["sample"]},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "rtfx": v / FRAME_RATE, "wall_s": time.perf_counter() - t_all}
    print(json.dumps(line), flush=True)


def run_ours(args) -> None:
    import torch
    import torch.distributed as dist
    from dia_tts_prune_b200 import synthetic as SY
    from dia_tts_prune_b200.engine import launch_count

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.set_num_threads(max(1, (os.cpu_count() or 1) // world))

    dia, cfg = build_cpu_model(WEIGHT_SEED, suppress_eos=True)
    if args.variant != "dense":
        # BASELINE.json configs[3]: the weight producers of dia/pruning_utils.py (+ the 2:4 mask) on the same model
        import torch.nn.utils.prune as prune
        from dia_tts_prune_b200 import pruning_utils as PU
        if args.variant == "mlp-pruned":
            for layer in dia.model.decoder.layers:
                prune.ln_structured(layer.mlp.wo, "weight", amount=args.prune_amount, n=2, dim=0)
        elif args.variant == "structured":
            PU.apply_structured_pruning(dia.model, args.prune_amount, dim=0, n=2)
            with torch.no_grad():        # keep the zeroed channel-0 EOS column from coming back as the only live one
                dia.model.decoder.logits_dense.weight[:, 0, cfg.data.audio_eos_value] = 0.0
        else:
            PU.apply_2to4_pruning(dia.model.decoder)
        PU.make_pruning_permanent(dia.model)
    sd_cpu = None
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        sd_cpu = {k: v.detach().clone() for k, v in dia.model.named_parameters()}
    # configs[1]: bf16 weights.  Values are bf16-representable, so this cast is exact.
    SY.cast_dense_kernels_(dia.model, torch.bfloat16)
    dia.compute_dtype = torch.bfloat16
    dia.device = dev
    dia.model.to(dev)
    dia.model.eval()
    text = SY.DEFAULT_TRANSCRIPT          # the same transcript on every rank and at every N: like-for-like weak scaling
    sampling = dict(cfg_scale=3.0, temperature=1.3, top_p=0.95, top_k=35)

    # ---- resident inputs for the device-timed loop --------------------------------------------------
    eff = dia._effective_text(text, None)
    with torch.inference_mode():
        dec_state, dec_out = dia._prepare_generation(eff, None, False)
    pristine = dec_out.generated_tokens.clone()
    text_len = dec_state.text_len
    eng_width = dia.model.decoder.engine().n_hidden           # < n_hidden when dead MLP neurons were dropped
    w_scale = 0.5625 if args.variant == "2to4" else 1.0
    live_bytes = dia.model.decoder.engine().weight_stream_bytes if args.variant == "structured" else None

    def one_generation(profile=None) -> int:
        for c in dec_state.self_attn_cache:
            c.current_idx = 0
        with torch.inference_mode():
            dec_out.generated_tokens.copy_(pristine)
            last = dia._run_loop(dec_state, dec_out, MAX_TOKENS, sampling["cfg_scale"], sampling["temperature"],
                                 sampling["top_p"], sampling["top_k"], 1234, False, profile=profile)
        return last + 1 - (dec_out.prefill_step - 1)        # loop iterations executed (the loop ends by `break`)

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        one_generation()
    sampler = ClockSampler(local)
    sync_all()
    sampler.start()
    launches0 = launch_count()
    profile: list = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    steps_run = 0
    for _ in range(args.steps):
        steps_run += one_generation(profile)
    e1.record()
    sync_all()
    launches = launch_count() - launches0
    dev_ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    frames_dev = steps_run                                     # one decode step = one frame (9 codes)

    # roofline of the step kernel: per-launch algorithmic bytes / event-timed duration
    k_bytes = k_ms = 0.0
    for (a, b, slot0, n) in profile:
        k_ms += a.elapsed_time(b)
        k_bytes += sum(algorithmic_bytes(cfg, slot0 + i, text_len, n_hidden=eng_width, weight_scale=w_scale, dense_bytes=live_bytes)
                       for i in range(n))
    peak, peak_src = measured_peak()
    achieved = k_bytes / (k_ms * 1e-3) / 1e9

    # ---- end to end through the public API, host buffers in and out ---------------------------------------
    def one_e2e():
        codes = dia.generate(text, max_tokens=MAX_TOKENS, seed=1234, output="codes", **{
            "cfg_scale": sampling["cfg_scale"], "temperature": sampling["temperature"], "top_p": sampling["top_p"],
            "cfg_filter_top_k": sampling["top_k"]})
        host = codes.cpu()
        return dia.last_stats["steps"], host
    for _ in range(min(args.warmup, 2)):
        one_e2e()
    sync_all()
    t0 = time.perf_counter()
    e2e_steps, d2h = 0, 0
    for _ in range(args.steps):
        n, host = one_e2e()
        e2e_steps += n
        d2h = host.numel() * host.element_size()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        dist.barrier()
    h2d = 2 * cfg.data.text_length * 8 + 4                     # [2, text_length] int64 text tokens (+ seed)

    # ---- reduce over ranks: max time, sum frames ------------------------------------------------------------
    t = torch.tensor([dev_ms, e2e_s, float(frames_dev), float(e2e_steps), k_ms, k_bytes, float(launches)],
                     dtype=torch.float64, device=dev)
    per_rank_ms = [dev_ms / args.steps]
    if world > 1:
        tmax, tsum = t.clone(), t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        every = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(every, t)
        per_rank_ms = [x[0].item() / args.steps for x in every]
        dev_ms, e2e_s = tmax[0].item(), tmax[1].item()
        frames_dev, e2e_steps, launches = tsum[2].item(), tsum[3].item(), tsum[6].item()
        k_ms, k_bytes = tsum[4].item() / world, tsum[5].item() / world            # roofline: mean over the replicas
        achieved = k_bytes / (k_ms * 1e-3) / 1e9
    value = frames_dev / (dev_ms * 1e-3)
    e2e_value = e2e_steps / e2e_s

    # ---- BASELINE.json configs[4] as a second field of the same line: 64 transcripts sharded over the ranks --------------------
    batch_field = None
    if args.batch_utterances > 0 and args.variant == "dense":
        from dia_tts_prune_b200 import replicas
        texts = replicas.shard([SY.synthetic_transcript(i) for i in range(args.batch_utterances)], world, rank)
        if texts:
            dia.generate_batch(texts[: args.per_launch], max_tokens=64, seed=1, max_utterances=args.per_launch)   # warm-up: repack
        sync_all()
        r = batch_workload(dia, cfg, texts, MAX_TOKENS, args.per_launch, dev) if texts else {"frames": 0, "secs": 0.0}
        frames_b, secs_b = replicas.reduce_throughput(r["frames"], r["secs"], device=dev)
        if rank == 0 and secs_b > 0:
            vb = frames_b / secs_b
            batch_field = {"workload": f"BASELINE configs[4]: {args.batch_utterances} synthetic transcripts (60-200 bytes) x "
                                       f"{MAX_TOKENS} tokens sharded round-robin over {world} GPU(s), {args.per_launch} "
                                       "utterances per kernel launch, end to end through Dia.generate_batch",
                           "value": vb, "unit": UNIT, "scaling": "strong", "frames": frames_b, "wall_s": secs_b,
                           "per_gpu_frames_per_s": vb / world, "n_per_launch": r.get("n_per_launch"),
                           "roofline_frames_per_s_per_gpu": r.get("roofline_frames_per_s_per_gpu"),
                           "frac_of_roofline": vb / world / r["roofline_frames_per_s_per_gpu"]
                           if r.get("roofline_frames_per_s_per_gpu") else None,
                           "rtfx": vb / FRAME_RATE}

    cpu_base = None
    if sd_cpu is not None:
        cpu_base = cpu_port_frames_per_s(sd_cpu, cfg, SY.DEFAULT_TRANSCRIPT, args.cpu_decode_steps, os.cpu_count() or 1,
                                         lean_steps=args.cpu_decode_steps)

    if rank == 0:
        # DRAM bytes per launch from the committed `ncu --set full` capture (per decode STEP there, scaled to the mean
        # number of steps of the launches timed here, so that it compares with algorithmic_bytes_per_launch)
        traffic, traffic_src = None, None
        steps_per_launch = frames_dev / world / max(1, len(profile))
        try:
            with open(os.path.join(REPO, "profiles", "step_kernel_traffic.json")) as f:
                tj = json.load(f)
            if args.variant == "dense":
                traffic = tj["dram_bytes_per_step"] * steps_per_launch
                traffic_src = tj.get("source")
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": dict(workload_config(text_len, world, args.variant, args.prune_amount),
                           **({"engine_mlp_width": eng_width} if args.variant != "dense" else {})),
            "decode_ms_per_frame": dev_ms / (frames_dev / world),
            "rtfx": value / world / FRAME_RATE,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1000.0 * e2e_s / args.steps},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "kernel": "dia_step_kernel",
                         "launches_timed": len(profile), "steps_per_launch": steps_per_launch,
                         "traffic_source": traffic_src,
                         "algorithmic_bytes_per_launch": k_bytes / max(1, len(profile)),
                         "algorithmic_bytes_per_step": k_bytes / max(1.0, frames_dev / world)},
            "per_rank_ms_per_step": per_rank_ms,
            "cpu_baseline": cpu_base,
            "clocks": clocks,
            "batch": batch_field,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def batch_workload(dia, cfg, texts, max_tokens: int, per_launch: int, dev) -> dict:
    """BASELINE.json configs[4] on ONE rank: its share of the transcripts through ``Dia.generate_batch`` (``per_launch``
    utterances per kernel launch: 2N rows share one pass over the weights), host text in, host codes out.  Returns frames,
    wall seconds and the roofline of this shape (bytes = W + N * 36 864 * e * (t + 2 Lt) per step of N
    frames)."""
    import torch
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    outs = dia.generate_batch(texts, max_tokens=max_tokens, seed=1234, max_utterances=per_launch)
    host = [o.cpu() for o in outs]
    torch.cuda.synchronize(dev)
    secs = time.perf_counter() - t0
    st = dia.last_stats
    frames = st["steps"]
    # roofline at this N: the mean context of a full generation, the mean text length of the transcripts
    lt = sum(len(dia._effective_text(t, None).encode("utf-8").replace(b"[S1]", b"\x01").replace(b"[S2]", b"\x02"))
             for t in texts) / max(1, len(texts))
    n = min(per_launch, len(texts))
    w_bytes = algorithmic_bytes(cfg, 0, 0) - cfg.model.decoder.n_layer * 2 * 2 * cfg.model.decoder.kv_heads * 128 * 4
    kv_per_utt = algorithmic_bytes(cfg, max_tokens // 2, int(lt)) - w_bytes
    peak, _ = measured_peak()
    roof = n / ((w_bytes + n * kv_per_utt) / (peak * 1e9))            # frames/s per GPU at the copy peak
    return {"frames": frames, "secs": secs, "loop_s": st["loop_s"], "prepare_s": st["prepare_s"], "n_per_launch": n,
            "roofline_frames_per_s_per_gpu": roof, "d2h_bytes": sum(h.numel() * h.element_size() for h in host)}


def run_batch(args) -> None:
    """BASELINE.json configs[4]: `--utterances M` synthetic transcripts sharded round-robin over the ranks (replicas, no
    data-path collective); every rank decodes its share `--per-launch` utterances at a time in one kernel.  Prints one
    JSON line with the whole-job throughput (sum of frames over ranks / slowest rank's wall time).  Not the headline
    line: the driver's `python bench.py` (no --utterances) stays configs[1] and carries this as its "batch" field."""
    import torch
    import torch.distributed as dist
    from dia_tts_prune_b200 import replicas, synthetic as SY

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    dia, cfg = build_cpu_model(WEIGHT_SEED, suppress_eos=True)
    SY.cast_dense_kernels_(dia.model, torch.bfloat16)
    dia.compute_dtype = torch.bfloat16
    dia.device = dev
    dia.model.to(dev)
    dia.model.eval()
    texts = [SY.synthetic_transcript(i) for i in range(args.utterances)]
    mine = replicas.shard(texts, world, rank)
    dia.generate_batch(mine[: args.per_launch], max_tokens=min(args.max_tokens, 64), seed=1, max_utterances=args.per_launch)
    if world > 1:
        dist.barrier()
    r = batch_workload(dia, cfg, mine, args.max_tokens, args.per_launch, dev)
    frames_all, secs_max = replicas.reduce_throughput(r["frames"], r["secs"], device=dev)
    if rank == 0:
        value = frames_all / secs_max
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "higher_is_better": True,
            "scaling": "strong", "data": "synthetic", "dtype": "bf16",
            "config": {"workload": f"{args.utterances} independent synthetic transcripts (60-200 bytes) sharded round-robin "
                                   f"over {world} GPU(s), {args.per_launch} utterances per kernel launch, max_tokens "
                                   f"{args.max_tokens} each, CFG batch 2 per utterance, default sampling; end to end "
                                   "through Dia.generate_batch (host text in, host codes out)",
                       "parallelism": f"replicas x{world}, no data-path collective"},
            "frames": frames_all, "wall_s": secs_max, "rtfx": value / FRAME_RATE,
            "per_gpu_frames_per_s": value / world,
            "roofline_frames_per_s_per_gpu": r["roofline_frames_per_s_per_gpu"],
            "frac_of_roofline": value / world / r["roofline_frames_per_s_per_gpu"]}), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-decode-steps", type=int, default=12, help="decode steps of the cpu_baseline sample")
    ap.add_argument("--ref-decode-steps", type=int, default=4, help="decode steps per reference-arm bench step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--variant", default="dense", choices=sorted(VARIANTS),
                    help="dense = BASELINE.json configs[1] (the headline); the others are configs[3]")
    ap.add_argument("--prune-amount", type=float, default=0.5, help="mlp-pruned: fraction of hidden neurons removed")
    ap.add_argument("--utterances", type=int, default=0,
                    help="BASELINE.json configs[4]: run this many synthetic transcripts sharded over the GPUs instead of "
                         "the headline single-transcript workload")
    ap.add_argument("--max-tokens", type=int, default=MAX_TOKENS, help="--utterances mode: frames per utterance")
    ap.add_argument("--per-launch", type=int, default=8, help="utterances decoded together in one kernel launch (1..8)")
    ap.add_argument("--batch-utterances", type=int, default=64,
                    help="headline mode: also run configs[4] with this many transcripts (0 = skip) and print it as `batch`")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        import socket
        s = socket.socket()
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
        s.close()
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    if args.utterances > 0:
        run_batch(args)
        return
    run_ours(args)


if __name__ == "__main__":
    main()
