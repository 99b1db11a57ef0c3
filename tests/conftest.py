import json
import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)
GOLD = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on a B200 with `-m gpu`)")


@pytest.fixture(scope="session")
def gold_tiny():
    return np.load(os.path.join(GOLD, "tiny_seed7.npz"))


@pytest.fixture(scope="session")
def gold_full():
    return np.load(os.path.join(GOLD, "dia16b_seed5_greedy.npz"))


@pytest.fixture(scope="session")
def gold_clone():
    return np.load(os.path.join(GOLD, "dia16b_seed5_clone861.npz"))


@pytest.fixture(scope="session")
def gold_delay():
    with open(os.path.join(GOLD, "delay_known_answer.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def gold_sampling():
    with open(os.path.join(GOLD, "sampling_known_answer.json")) as f:
        return json.load(f)


def build_dia(cfg, seed, device=None, bf16=False):
    """Product Dia with the shared synthetic weights; returns (dia, fp32 CPU state dict)."""
    import torch
    from dia_tts_prune_b200 import synthetic as SY
    from dia_tts_prune_b200.model import Dia
    dia = Dia(cfg, "float32", torch.device("cpu"))
    SY.init_synthetic_(dia.model.named_parameters(), seed)
    sd = {k: v.detach().clone() for k, v in dia.model.named_parameters()}
    if device is not None:
        if bf16:
            SY.cast_dense_kernels_(dia.model, torch.bfloat16)
            dia.compute_dtype = torch.bfloat16
        dia.device = torch.device(device)
        dia.model.to(dia.device)
    dia.model.eval()
    return dia, sd


@pytest.fixture(scope="session")
def tiny_gpu():
    import torch
    from dia_tts_prune_b200.config import tiny_config
    dia, sd = build_dia(tiny_config(), 7, "cuda:0")
    yield dia, sd
    torch.cuda.synchronize()


@pytest.fixture(scope="session")
def full_gpu():
    """Dia-1.6B, seed 5 (the seed screened for a healthy greedy margin), bf16 weights on cuda:0."""
    import torch
    from dia_tts_prune_b200.config import dia_1_6b_config
    dia, sd = build_dia(dia_1_6b_config(), 5, "cuda:0", bf16=True)
    yield dia, sd
    torch.cuda.synchronize()
