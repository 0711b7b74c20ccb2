"""Synthetic OCM-like inputs for benchmarks and demos: a dark, right-skewed, spatially correlated gray field with
thin bright "fibres", quantised to k/255 like ``ToTensor()`` of a uint8 image.  Deterministic per seed.  (The test oracle carries
its own copy of the recipe; tests/test_oracle.py checks that the two stay identical.)"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def synthetic_gray(size: int, seed: int = 1234, batch: int = 1) -> torch.Tensor:
    """[B, 1, S, S] fp32 in [0, 1], values k / 255."""
    g = torch.Generator().manual_seed(seed)
    low = torch.rand(batch, 1, max(size // 8, 2), max(size // 8, 2), generator=g)
    up = F.interpolate(low, size=(size, size), mode="bicubic", align_corners=False)
    noise = torch.rand(batch, 1, size, size, generator=g) ** 2
    yy = torch.arange(size).view(1, 1, size, 1).float()
    xx = torch.arange(size).view(1, 1, 1, size).float()
    fibers = 0.10 * (torch.sin(0.11 * xx + 0.07 * yy) > 0.85).float()
    img = (0.12 * torch.exp(1.5 * (up - 0.5)) + 0.08 * noise + fibers).clamp(0, 1)
    return torch.floor(img * 255.0) / 255.0


def synthetic_tile(size: int = 224, seed: int = 1234, batch: int = 1) -> torch.Tensor:
    """[B, 3, S, S] fp32 with R = G = B (real OCM images are gray)."""
    return synthetic_gray(size, seed, batch).expand(-1, 3, -1, -1).contiguous()


def synthetic_mosaic_u8(size: int, seed: int = 4321) -> np.ndarray:
    """[S, S] uint8 gray mosaic."""
    g = synthetic_gray(size, seed, 1)[0, 0]
    return (g * 255.0).round().to(torch.uint8).numpy()


def random_masks(rng: np.random.RandomState, batch: int, input_size: int = 224, mask_patch_size: int = 16, model_patch_size: int = 8,
                 mask_ratio: float = 0.5) -> torch.Tensor:
    """[B, S/p, S/p] int64 SimMIM masks with the MaskGenerator recipe (SSS/data.py:163-186) from an explicit RandomState."""
    rand_size = input_size // mask_patch_size
    scale = mask_patch_size // model_patch_size
    count = rand_size ** 2
    mask_count = int(np.ceil(count * mask_ratio))
    out = []
    for _ in range(batch):
        idx = rng.permutation(count)[:mask_count]
        m = np.zeros(count, dtype=int)
        m[idx] = 1
        out.append(m.reshape(rand_size, rand_size).repeat(scale, axis=0).repeat(scale, axis=1))
    return torch.from_numpy(np.stack(out))
