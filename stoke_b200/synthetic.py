# -*- coding: utf-8 -*-
"""Synthetic workload recipes for BASELINE.json's configs (shared by bench.py, the tests and the golden generator).

No network: models are random-init, data is drawn from seeded generators.  Pure torch/numpy -- importing this module does
not load the CUDA library.
"""
from typing import Iterator, Tuple

import numpy as np
import torch

# ---- configs[0]: BasicNN + BCEWithLogitsLoss + Adam (README.md:100-137 of the reference leaves BasicNN as ``pass``;
# SURVEY.md section 8 fixes it as Linear(128,256)-ReLU-Linear(256,256)-ReLU-Linear(256,1) = 99,073 parameters) ----------
CFG1_BATCH = 32
CFG1_ACCUM = 2
CFG1_OPT_STEPS = 50
CFG1_ADAM = {"lr": 1e-3, "betas": (0.9, 0.98), "eps": 1e-9}


def basic_nn(seed: int = 0) -> torch.nn.Module:
    torch.manual_seed(seed)
    return torch.nn.Sequential(
        torch.nn.Linear(128, 256), torch.nn.ReLU(), torch.nn.Linear(256, 256), torch.nn.ReLU(), torch.nn.Linear(256, 1)
    )


def cfg1_batches(n_micro: int, seed: int = 1) -> Iterator[Tuple[torch.Tensor, torch.Tensor]]:
    g = torch.Generator()
    g.manual_seed(seed)
    for _ in range(n_micro):
        x = torch.randn(CFG1_BATCH, 128, generator=g)
        y = (torch.rand(CFG1_BATCH, 1, generator=g) > 0.5).float()
        yield x, y


# ---- configs[1..2]: ResNet-50 synthetic 3x224x224 ---------------------------------------------------------------------
def resnet50(seed: int = 0) -> torch.nn.Module:
    import torchvision

    torch.manual_seed(seed)
    return torchvision.models.resnet50(num_classes=1000)


def resnet50_batch(batch: int, rank: int = 0, size: int = 224) -> Tuple[torch.Tensor, torch.Tensor]:
    g = torch.Generator()
    g.manual_seed(1000 + rank)
    x = torch.randn(batch, 3, size, size, generator=g)
    y = torch.randint(0, 1000, (batch,), generator=g)
    return x, y


# ---- configs[3]: sampler workload (lengths -> stable argsort) ---------------------------------------------------------
def sampler_lengths(n: int, lo: int = 1, hi: int = 513, seed: int = 0) -> np.ndarray:
    return np.random.default_rng(seed).integers(lo, hi, size=n)


def sampler_sorted_idx(n: int, lo: int = 1, hi: int = 513, seed: int = 0) -> np.ndarray:
    return np.argsort(sampler_lengths(n, lo, hi, seed), kind="stable")


# ---- configs[4] / gradient-injection: seeded per-rank gradient streams -------------------------------------------------
def injected_grad(n: int, rank: int, step: int, dtype=torch.bfloat16, scale: float = 1.0) -> torch.Tensor:
    """Rank ``rank``'s flat gradient for step ``step`` (values exactly representable in ``dtype``)."""
    g = torch.Generator()
    g.manual_seed(2000 + 7919 * step + rank)
    v = torch.randn(n, generator=g) * scale
    return v.to(dtype)
