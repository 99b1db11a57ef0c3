"""``Dia`` inference driver (drop-in for the reference's ``dia/model.py``).

Public surface kept: ``Dia(config, compute_dtype, device)``, ``Dia.from_local``,
``Dia.from_pretrained``, ``Dia.generate`` (same keyword arguments and defaults),
``Dia.save_audio``, ``Dia.load_audio``, ``ComputeDtype``, ``DEFAULT_SAMPLE_RATE``
(dia/model.py:21,85-98,102-236,546-595,631-846).

What is different underneath: the per-frame loop of ``generate`` (dia/model.py:748-815) runs as a
device-resident state machine inside the persistent step kernel - embedding gather-sum, 18 decoder
layers, logits head, CFG combine, masks, argmax or top-k/top-p sampling, the EOS countdown and
the BOS-masked write-back - so a run of decode steps costs one kernel launch and no host round trip.
``_decoder_step`` remains available as the single-step boundary.

Declared deviations from the reference (all in its favour or outside the kernel scope):
  * the shipped loop raises before the first token; the intended semantics,
    including every quirk of Appendix C, are implemented;
  * DAC (third-party codec) is loaded lazily and only if installed; ``generate(..., output="codes")``
    returns the DAC-ready code tensor without it;
  * ``use_torch_compile`` is accepted and ignored (there is no tracing compiler on this path);
  * sampling draws come from a Philox stream keyed by ``seed`` (or by the torch global RNG when
    ``seed`` is None), not from ``torch.multinomial``'s stream: same distribution, different draws.
"""

from __future__ import annotations

import random
import time
from enum import Enum
from pathlib import Path
from typing import Optional

import numpy as np
import torch

from . import audio as _audio
from .audio import decode  # noqa: F401  (re-export, as in the reference)
from .config import DiaConfig
from .layers import DiaModel
from .state import DecoderInferenceState, DecoderOutput, EncoderInferenceState

try:
    import dac  # type: ignore
except Exception:  # the codec is optional here
    dac = None
try:
    import soundfile as sf  # type: ignore
except Exception:
    sf = None

DEFAULT_SAMPLE_RATE = 44100
_STEPS_PER_LAUNCH = 128


def _get_default_device() -> torch.device:
    if torch.cuda.is_available():
        return torch.device("cuda")
    return torch.device("cpu")


class ComputeDtype(str, Enum):
    FLOAT32 = "float32"
    FLOAT16 = "float16"
    BFLOAT16 = "bfloat16"

    def to_dtype(self) -> torch.dtype:
        return {"float32": torch.float32, "float16": torch.float16, "bfloat16": torch.bfloat16}[self.value]


def _sample_next_token(logits_BCxV: torch.Tensor, temperature: float, top_p: float,
                       cfg_filter_top_k: int | None = None) -> torch.Tensor:
    """Module-level helper kept for API compatibility (dia/model.py:32-82): samples one token per row
    of already-guided logits.  The generate loop does NOT call this - it samples in the kernel."""
    if temperature == 0.0:
        return torch.argmax(logits_BCxV, dim=-1)
    x = logits_BCxV / temperature
    if cfg_filter_top_k is not None and cfg_filter_top_k > 0:
        kth = torch.topk(x, k=cfg_filter_top_k, dim=-1).values[..., -1:]
        x = x.masked_fill(x < kth, -torch.inf)
    if top_p < 1.0:
        sp, si = torch.sort(torch.softmax(x, dim=-1), dim=-1, descending=True)
        drop = torch.cumsum(sp, dim=-1) > top_p
        drop = torch.cat([torch.zeros_like(drop[..., :1]), drop[..., :-1]], dim=-1)
        x = x.masked_fill(torch.zeros_like(drop).scatter(-1, si, drop), -torch.inf)
    return torch.multinomial(torch.softmax(x, dim=-1), num_samples=1).squeeze(-1)


class Dia:
    def __init__(self, config: DiaConfig, compute_dtype: str | ComputeDtype = ComputeDtype.FLOAT32,
                 device: torch.device | None = None):
        self.config = config
        self.device = torch.device(device) if device is not None else _get_default_device()
        if isinstance(compute_dtype, str):
            compute_dtype = ComputeDtype(compute_dtype)
        self.compute_dtype = compute_dtype.to_dtype()
        if self.device.type == "cpu" and self.compute_dtype != torch.float32:
            print(f"Warning: CPU device selected, overriding compute_dtype to float32 (was {compute_dtype.value}).")
            self.compute_dtype = torch.float32
        # float16 (the default of the reference's cli.py / app.py) is accepted: the kernels keep activations in fp32
        # and stream weights as bf16, so float16 kernels are widened and rounded once (layers.canonicalize_dense_kernel_)
        self.model: DiaModel = DiaModel(config, self.compute_dtype)
        self._draws = 0                                 # RNG draw counter of the single-step boundary (_decoder_step)
        self._step_seed: int | None = None
        self.dac_model = None
        self.last_codes: torch.Tensor | None = None     # raw generated rows of the last generate() call
        self.last_stats: dict = {}
        self.batch_min_utterances = 2                   # generate_batch: smaller groups run through the single-utterance kernel
        # encode / project only the text bytes the decoder can observe (set False for the reference's full tensors)
        self.live_text_only = True

    # ---- loading ---------------------------------------------------------------------------------
    @classmethod
    def from_local(cls, config_path: str, checkpoint_path: str,
                   compute_dtype: str | ComputeDtype = ComputeDtype.FLOAT32,
                   device: torch.device | None = None) -> "Dia":
        config = DiaConfig.load(config_path)
        if config is None:
            raise FileNotFoundError(f"Config file not found at {config_path}")
        dia = cls(config, compute_dtype, device)
        try:
            sd = torch.load(checkpoint_path, map_location="cpu")
            sd = {k: v for k, v in sd.items() if "lora_" not in k}
            missing, unexpected = dia.model.load_state_dict(sd, strict=False)
            if unexpected:
                print(f"Warning: Unexpected keys found in checkpoint: {unexpected}")
            if missing:
                print(f"Warning: Missing keys in checkpoint: {missing}")
        except FileNotFoundError:
            raise FileNotFoundError(f"Checkpoint file not found at {checkpoint_path}")
        except Exception as e:
            raise RuntimeError(f"Error loading checkpoint from {checkpoint_path}") from e
        dia.model.to(dia.device)
        dia.model.eval()
        dia._load_dac_model(required=False)
        return dia

    @classmethod
    def from_pretrained(cls, model_name: str = "nari-labs/Dia-1.6B",
                        compute_dtype: str | ComputeDtype = ComputeDtype.FLOAT32,
                        device: torch.device | None = None, **kwargs) -> "Dia":
        if isinstance(compute_dtype, str):
            compute_dtype = ComputeDtype(compute_dtype)
        loaded = DiaModel.from_pretrained(model_name, **kwargs)
        dia = cls(loaded.config, compute_dtype, device if device is not None else _get_default_device())
        dia.model = loaded
        dia.model.to(dia.device)
        dia.model.eval()
        dia._load_dac_model(required=False)
        return dia

    def _load_dac_model(self, required: bool = True):
        """Descript Audio Codec (third party).  Optional on this path: codes can be produced without it."""
        if dac is None or not hasattr(dac, "utils"):
            if required:
                raise RuntimeError("Failed to load DAC model: descript-audio-codec is not installed")
            return
        try:
            m = dac.DAC.load(dac.utils.download()).to(self.device)
            m.eval()
            self.dac_model = m
        except Exception as e:
            if required:
                raise RuntimeError(f"Failed to load DAC model: {e}") from e

    # ---- input preparation (dia/model.py:254-427) ----------------------------------------------------
    def _prepare_text_input(self, text: str) -> torch.Tensor:
        n, pad = self.config.data.text_length, self.config.data.text_pad_value
        toks = list(text.encode("utf-8").replace(b"[S1]", b"\x01").replace(b"[S2]", b"\x02"))
        if len(toks) > n:
            print(f"Warning: Input text truncated from {len(toks)} to {n} bytes.")
            toks = toks[:n]
        out = torch.full((1, n), pad, dtype=torch.long)
        out[0, :len(toks)] = torch.tensor(toks, dtype=torch.long)
        return out.to(self.device)

    def _prepare_audio_prompt(self, audio_prompt: torch.Tensor | None) -> tuple[torch.Tensor, int]:
        dt = self.config.data
        rows = [torch.full((1, dt.channels), dt.audio_bos_value, dtype=torch.int32, device=self.device)]
        prefill_step = 1
        if audio_prompt is not None:
            if audio_prompt.ndim == 3 and audio_prompt.shape[0] == 1:
                audio_prompt = audio_prompt.squeeze(0)
            if audio_prompt.ndim != 2:
                raise ValueError(f"Unexpected audio_prompt shape: {audio_prompt.shape}. Expected [T, C] or [1, T, C].")
            prefill_step += audio_prompt.shape[0]
            rows.append(audio_prompt.to(device=self.device, dtype=torch.int32))
        # the delay tail is PAD, not -1 (dia/model.py:330-333) - which is why the first 14 predictions
        # after a prompt are discarded by the masked write-back
        rows.append(torch.full((max(dt.delay_pattern), dt.channels), dt.audio_pad_value, dtype=torch.int32,
                               device=self.device))
        grid = torch.cat(rows, dim=0)
        delayed = _audio.delay_apply(grid.unsqueeze(0), dt.delay_pattern, dt.audio_pad_value, dt.audio_bos_value)
        return delayed.squeeze(0), prefill_step

    def _prepare_generation(self, text: str, audio_prompt: str | torch.Tensor | None, verbose: bool):
        cond = self._prepare_text_input(text)
        enc_input = torch.cat([torch.full_like(cond, self.config.data.text_pad_value), cond], dim=0)
        if isinstance(audio_prompt, str):
            audio_prompt = self.load_audio(audio_prompt)
        delayed, prefill_step = self._prepare_audio_prompt(audio_prompt if isinstance(audio_prompt, torch.Tensor) else None)
        if verbose:
            print(f"generate: Text tokens shape: {enc_input.shape}")
            print(f"generate: Prefill audio steps: {prefill_step}")
        enc_state = EncoderInferenceState.new(self.config, enc_input)
        self.model.eval()
        with torch.inference_mode():
            # valid text bytes form a prefix (pad = 0 never occurs inside utf-8 text); known on the host
            n_valid = int((cond[0] != self.config.data.text_pad_value).sum().item())
            if not bool((cond[0, :n_valid] != self.config.data.text_pad_value).all().item()):
                raise NotImplementedError("text with embedded pad bytes is not supported by the cross-attention kernel")
            if self.live_text_only and n_valid > 0:
                # Only the n_valid text bytes of the conditional row are ever observed by the decoder: the
                # unconditional row and the pad positions are masked everywhere, and the
                # encoder never mixes pad and non-pad tokens (dia/state.py:24-31).  Encode and project just those.
                live = cond[:, :n_valid]
                es_live = EncoderInferenceState(
                    max_seq_len=n_valid, device=live.device,
                    positions=torch.arange(n_valid, dtype=torch.float32, device=live.device)[None, :],
                    padding_mask=torch.ones_like(live, dtype=torch.bool), attn_mask=None, valid_lens=[n_valid])
                enc_live = self.model.encoder(live, es_live)
                enc_out = torch.zeros((2, cond.shape[1], enc_live.shape[-1]), dtype=enc_live.dtype, device=live.device)
                enc_out[1, :n_valid] = enc_live[0]
                cross = self.model.decoder.precompute_cross_attn_cache_live(enc_live, cond.shape[1])
            else:
                enc_out = self.model.encoder(enc_input, enc_state)
                cross = self.model.decoder.precompute_cross_attn_cache(enc_out, enc_state.positions)
            dec_state = DecoderInferenceState.new(self.config, enc_state, enc_out, cross, self.compute_dtype)
            dec_state.text_len = n_valid
            dec_output = DecoderOutput.new(self.config, self.device)
            dec_output.prefill(delayed, prefill_step)
            if prefill_step > 1:
                dec_state.prepare_step(0, prefill_step - 1)
                toks = dec_output.get_tokens_at(0, prefill_step - 1).unsqueeze(0).expand(2, -1, -1)
                self.model.decoder.forward(toks, dec_state, want_logits=False)
        return dec_state, dec_output

    # ---- single-step boundary (dia/model.py:429-488) ------------------------------------------------------
    def _decoder_step(self, tokens_Bx1xC: torch.Tensor, dec_state: DecoderInferenceState, cfg_scale: float,
                      temperature: float, top_p: float, cfg_filter_top_k: int | None, seed: int | None = None,
                      draw: int | None = None) -> torch.Tensor:
        """One frame (dia/model.py:429-488).  A reference-style caller loops this with the six positional arguments:
        the Philox stream is then keyed once from the torch global RNG (so ``torch.manual_seed`` governs it, as it
        governs ``torch.multinomial`` in the reference) and every call consumes the next draw index."""
        logits = self.model.decoder.decode_step(tokens_Bx1xC, dec_state)          # [2, 1, C, V] fp32
        eng = self.model.decoder.engine()
        if seed is None:
            if self._step_seed is None:
                self._step_seed = int(torch.randint(0, 2 ** 62, (1,)).item())
            seed = self._step_seed
        if draw is None:
            draw = self._draws
            self._draws += 1
        return eng.head_sample(logits[:, -1], cfg_scale, temperature, top_p, cfg_filter_top_k, seed, draw).to(torch.int64)

    # ---- output (dia/model.py:490-544) -----------------------------------------------------------------------
    def _finalize_codes(self, generated_codes: torch.Tensor) -> torch.Tensor:
        dt = self.config.data
        size = self.dac_model.vq_config.codebook_size if (self.dac_model is not None and hasattr(self.dac_model, "vq_config")) else 1024
        return _audio.finalize_codes(generated_codes, dt.delay_pattern, dt.audio_pad_value, size)

    def _generate_output(self, generated_codes: torch.Tensor) -> np.ndarray | None:
        if self.dac_model is None:
            raise RuntimeError("DAC model not loaded. Cannot decode audio.")
        if generated_codes is None or generated_codes.numel() == 0:
            print("Warning: No generated codes to decode.")
            return None
        codes = self._finalize_codes(generated_codes)
        try:
            with torch.inference_mode():
                wav = _audio.decode(self.dac_model, codes)
        except Exception as e:
            print(f"Error during DAC decoding: {e}")
            return None
        return wav.squeeze().cpu().numpy()

    def load_audio(self, audio_path: str) -> torch.Tensor:
        if self.dac_model is None:
            self._load_dac_model(required=True)
        try:
            import torchaudio
            wav, sr = torchaudio.load(audio_path)
            if wav.shape[0] > 1:
                wav = wav.mean(dim=0, keepdim=True)
            if sr != DEFAULT_SAMPLE_RATE:
                wav = torchaudio.functional.resample(wav, sr, DEFAULT_SAMPLE_RATE)
            wav = wav.to(self.device).unsqueeze(0)
            with torch.inference_mode():
                data = self.dac_model.preprocess(wav, DEFAULT_SAMPLE_RATE)
                _, codes, _, _, _ = self.dac_model.encode(data)
            return codes.squeeze(0).transpose(0, 1)
        except FileNotFoundError:
            raise FileNotFoundError(f"Audio file not found: {audio_path}")
        except Exception as e:
            raise RuntimeError(f"Error loading or encoding audio file {audio_path}: {e}") from e

    def save_audio(self, path: str, audio: np.ndarray, sample_rate: int = DEFAULT_SAMPLE_RATE):
        if audio is None:
            print("Warning: Cannot save None audio.")
            return
        try:
            if sf is None:
                raise RuntimeError("soundfile is not installed")
            Path(path).parent.mkdir(parents=True, exist_ok=True)
            if not np.issubdtype(audio.dtype, np.floating):
                audio = audio.astype(np.float32) / np.iinfo(audio.dtype).max
            sf.write(path, np.clip(audio, -1.0, 1.0), sample_rate)
        except Exception as e:
            print(f"Error saving audio to {path}: {e}")

    def load_adapter_weights(self, adapter_path: str, adapter_name: str = "default"):
        try:
            from peft import PeftModel  # noqa: F401
        except Exception as e:
            raise ImportError("PEFT library is required to load adapters. Install with `pip install peft`.") from e
        if not hasattr(self.model, "load_adapter"):
            raise RuntimeError("Model is not a PEFT model. Load adapters onto a model previously configured with PEFT.")
        self.model.load_adapter(adapter_path, adapter_name=adapter_name)
        self.model.set_adapter(adapter_name)
        self.model.decoder.invalidate_engine()

    # ---- generate (dia/model.py:631-846) ------------------------------------------------------------------------
    @staticmethod
    def _effective_text(text: str, audio_prompt_text: Optional[str]) -> str:
        t = audio_prompt_text.strip() + " " + text.strip() if audio_prompt_text else text.strip()
        s1, s2 = t.rfind("[S1]"), t.rfind("[S2]")
        if s1 > s2 and not t.endswith("[S2]"):
            t += " [S2]"
        elif s2 > s1 and not t.endswith("[S1]"):
            t += " [S1]"
        elif s1 == -1 and s2 == -1 and t:
            t += " [S2]"
        return t

    @torch.inference_mode()
    def generate(self, text: str, max_tokens: int | None = None, cfg_scale: float = 3.0, temperature: float = 1.3,
                 top_p: float = 0.95, use_torch_compile: bool = False, cfg_filter_top_k: int | None = 35,
                 audio_prompt: str | torch.Tensor | None = None, audio_prompt_text: Optional[str] = None,
                 seed: Optional[int] = None, verbose: bool = False, output: str = "audio"):
        """Text (+ optional voice prompt) -> waveform (``output="audio"``, needs DAC) or the DAC-ready
        code tensor int32 [1, C, T] (``output="codes"``).  Returns None on failure, like the reference."""
        if audio_prompt is not None and not audio_prompt_text:
            raise ValueError("`audio_prompt_text` is required when `audio_prompt` is provided.")
        if output not in ("audio", "codes"):
            raise ValueError("output must be 'audio' or 'codes'")
        if self.device.type != "cuda" or not torch.cuda.is_available():
            # loud, before the reference's catch-all: this path has no CPU implementation
            raise RuntimeError("Dia.generate needs the model on a CUDA device (sm_100a); there is no CPU fallback")
        if temperature < 0.0:
            raise ValueError("temperature must be >= 0")
        if seed is not None:
            torch.manual_seed(seed)
            np.random.seed(seed)
            random.seed(seed)
            self._step_seed, self._draws = None, 0
        rng_seed = int(seed) if seed is not None else int(torch.randint(0, 2 ** 62, (1,)).item())
        max_tokens = self.config.data.audio_length if max_tokens is None else max_tokens
        t_start = time.time()
        try:
            dec_state, dec_output = self._prepare_generation(self._effective_text(text, audio_prompt_text),
                                                             audio_prompt, verbose)
        except Exception as e:
            print(f"Error during preparation: {e}")
            import traceback
            traceback.print_exc()
            return None
        t_prep = time.time()
        try:
            dec_step = self._run_loop(dec_state, dec_output, max_tokens, cfg_scale, temperature, top_p,
                                      cfg_filter_top_k, rng_seed, verbose)
        except Exception as e:
            print(f"Error during generation loop: {e}")
            import traceback
            traceback.print_exc()
            return None
        if dec_output.prefill_step > dec_step + 1:
            print("Warning: No new tokens were generated after prefill.")
            return None
        codes = dec_output.generated_tokens[dec_output.prefill_step: dec_step + 1, :]
        self.last_codes = codes
        self.last_stats = {"prepare_s": t_prep - t_start, "loop_s": time.time() - t_prep,
                           "steps": dec_step + 1 - (dec_output.prefill_step - 1), "frames": int(codes.shape[0])}
        if verbose:
            s = self.last_stats
            print(f"generate: Total steps generated={s['frames']}, loop {s['loop_s']:.3f}s "
                  f"({s['steps'] / max(s['loop_s'], 1e-9):.1f} tokens/s), prepare {s['prepare_s']:.3f}s")
        if output == "codes":
            return self._finalize_codes(codes)
        try:
            return self._generate_output(codes)
        except Exception as e:
            print(f"Error during final decoding: {e}")
            return None

    # ---- N utterances per launch (no counterpart in the reference, whose batch is the CFG pair) ----
    @torch.inference_mode()
    def generate_batch(self, texts: list[str], max_tokens: int | None = None, cfg_scale: float = 3.0,
                       temperature: float = 1.3, top_p: float = 0.95, cfg_filter_top_k: int | None = 35,
                       seed: Optional[int] = None, max_utterances: int = 8, output: str = "codes", verbose: bool = False,
                       audio_prompts: Optional[list] = None, audio_prompt_texts: Optional[list] = None):
        """``generate`` for a list of independent transcripts, decoded ``max_utterances`` at a time in ONE kernel: the 2N
        CFG rows share a single pass over the weights per frame.  Every utterance has its own encoder pass, KV caches,
        token grid, EOS state and RNG stream (utterance i draws from seed + i), and yields exactly the rows
        ``generate(texts[i], ...)`` yields for greedy decoding.  ``audio_prompts`` / ``audio_prompt_texts`` (lists, entries
        may be None) are the voice-clone arguments of ``generate`` per utterance: prompts of different lengths start at
        different cache slots of the same launch.  Returns a list (codes int32 [1, C, T_i] or waveforms)."""
        if output not in ("audio", "codes"):
            raise ValueError("output must be 'audio' or 'codes'")
        if self.device.type != "cuda" or not torch.cuda.is_available():
            raise RuntimeError("Dia.generate_batch needs the model on a CUDA device (sm_100a); there is no CPU fallback")
        if temperature < 0.0:
            raise ValueError("temperature must be >= 0")
        audio_prompts = list(audio_prompts) if audio_prompts is not None else [None] * len(texts)
        audio_prompt_texts = list(audio_prompt_texts) if audio_prompt_texts is not None else [None] * len(texts)
        if len(audio_prompts) != len(texts) or len(audio_prompt_texts) != len(texts):
            raise ValueError("audio_prompts / audio_prompt_texts must have one entry per transcript")
        for ap, apt in zip(audio_prompts, audio_prompt_texts):
            if ap is not None and not apt:
                raise ValueError("`audio_prompt_text` is required when `audio_prompt` is provided.")
        if seed is not None:
            torch.manual_seed(seed)
        base_seed = int(seed) if seed is not None else int(torch.randint(0, 2 ** 31, (1,)).item())   # + i stays a numpy seed
        max_tokens = self.config.data.audio_length if max_tokens is None else max_tokens
        results: list = [None] * len(texts)
        stats = {"prepare_s": 0.0, "loop_s": 0.0, "steps": 0, "frames": 0, "launch_steps": 0}
        raw: list = [None] * len(texts)
        per = max(1, min(int(max_utterances), 8))
        for b0 in range(0, len(texts), per):
            idx = list(range(b0, min(len(texts), b0 + per)))
            t0 = time.time()
            if len(idx) < self.batch_min_utterances:
                # Below the measured crossover (tools/batch_bench.py at cache slot 1500: the batched step costs 1.20 ms for one
                # utterance, 1.26 for 2, 1.40 for 4 and 1.62 for 8; the single-utterance step 0.75 ms) a single utterance is
                # faster on its own kernel
                for i in idx:
                    res = self.generate(texts[i], max_tokens=max_tokens, cfg_scale=cfg_scale, temperature=temperature,
                                        top_p=top_p, cfg_filter_top_k=cfg_filter_top_k, seed=base_seed + i, output=output,
                                        audio_prompt=audio_prompts[i], audio_prompt_text=audio_prompt_texts[i])
                    results[i], raw[i] = res, self.last_codes
                    stats["prepare_s"] += self.last_stats.get("prepare_s", 0.0)
                    stats["loop_s"] += self.last_stats.get("loop_s", 0.0)
                    stats["steps"] += self.last_stats.get("steps", 0)
                    stats["frames"] += self.last_stats.get("frames", 0)
                continue
            eng = self.model.decoder.batch_engine(per)
            prepared = [self._prepare_generation(self._effective_text(texts[i], audio_prompt_texts[i]), audio_prompts[i], False)
                        for i in idx]
            for u, (st, out) in enumerate(prepared):
                for c in st.cross_attn_cache:
                    if c.k.dtype != torch.float32 or not c.k.is_contiguous():
                        c.k = c.k.to(torch.float32).contiguous()
                    if c.v.dtype != torch.float32 or not c.v.is_contiguous():
                        c.v = c.v.to(torch.float32).contiguous()
                eng.bind(u, st.self_attn_cache, st.cross_attn_cache, st.text_len)
            P = [out.prefill_step for _, out in prepared]
            slots = [st.self_attn_cache[0].current_idx for st, _ in prepared]
            eng.generate_begin([out.generated_tokens for _, out in prepared], P, slots, max_tokens, cfg_scale, temperature,
                               top_p, cfg_filter_top_k, [base_seed + i for i in idx])
            torch.cuda.synchronize(self.device)
            t1 = time.time()
            remaining = max(0, max_tokens - min(P))
            while remaining > 0:
                n = min(_STEPS_PER_LAUNCH, remaining)
                eng.generate_steps(n)
                remaining -= n
                stats["launch_steps"] += n
            sts = eng.status()
            t2 = time.time()
            stats["prepare_s"] += t1 - t0
            stats["loop_s"] += t2 - t1
            for u, i in enumerate(idx):
                st, out = prepared[u]
                for c in st.self_attn_cache:
                    c.current_idx = slots[u] + sts[u].steps_run
                codes = out.generated_tokens[out.prefill_step: sts[u].dec_step + 1, :]
                raw[i] = codes
                stats["steps"] += sts[u].steps_run
                stats["frames"] += int(codes.shape[0])
                results[i] = self._finalize_codes(codes) if output == "codes" else self._generate_output(codes)
            if verbose:
                print(f"generate_batch: utterances {idx[0]}..{idx[-1]}: prepare {t1 - t0:.3f}s, loop {t2 - t1:.3f}s")
        self.last_batch_codes = raw
        self.last_stats = stats
        return results

    def _run_loop(self, dec_state: DecoderInferenceState, dec_output: DecoderOutput, max_tokens: int, cfg_scale: float,
                  temperature: float, top_p: float, top_k: int, seed: int, verbose: bool,
                  profile: list | None = None) -> int:
        """The while-loop of dia/model.py:748-815, executed on the device in blocks of steps.
        ``profile``: optional list receiving (start_event, end_event, first_slot, n_steps) per launch."""
        eng = self.model.decoder._engine_for(dec_state)
        first_slot = dec_state.self_attn_cache[0].current_idx
        P = dec_output.prefill_step
        eng.generate_begin(dec_output.generated_tokens, P, first_slot, max_tokens, cfg_scale, temperature, top_p,
                           top_k, seed)
        remaining = max(0, max_tokens - P)
        # all launches are queued back to back; a launch that starts after the loop finished is a no-op
        slot = first_slot
        while remaining > 0:
            n = min(_STEPS_PER_LAUNCH, remaining)
            if profile is not None:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                eng.generate_steps(n)
                e1.record()
                profile.append((e0, e1, slot, n))
            else:
                eng.generate_steps(n)
            remaining -= n
            slot += n
        st = eng.status()                                  # the only host <-> device sync of the loop
        for c in dec_state.self_attn_cache:
            c.current_idx = first_slot + st.steps_run
        if verbose:
            print(f"generate: device loop ran {st.steps_run} steps, dec_step={st.dec_step}, "
                  f"eos_detected={bool(st.eos_detected)}")
        return st.dec_step
