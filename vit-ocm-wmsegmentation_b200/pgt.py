"""Pseudo-ground-truth mask generation of the reference's ``PGT.py`` on B200.

``PGT.train`` / ``PGT.evaluate`` (SSS/PGT.py:50-97, :100-146) build, for every image of every batch of every epoch, a
pseudo mask out of the frozen encoder's CLS attention: ``get_intermediate_feat`` -> ``compute_attention`` -> mean over
(all, or a random subset of) heads -> resize pair -> ``utils.threshold`` -> ``y[i] = output / 255`` -- one image at a time with
a ``.cpu()`` round trip each.  ``pseudo_masks`` does the whole batch on the device: one ``vitocm_forward_cls_attn`` for the
batch, the head reduction, bilinear upsampling, image/attention blend, Otsu and binarisation in ``vitocm_tile_threshold``; the
result never leaves HBM and feeds the student's loss directly."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from ._lib import check, cur_stream, ptr
from .utils import head_mean_maps


def select_heads(batch: int, num_heads: int, rng=np.random) -> torch.Tensor:
    """The ``rand = True`` branch of SSS/PGT.py:66-73, drawn from numpy's RNG in the reference's order (per image: ``randint(1, 7)``
    then ``choice(heads, size, replace=False)``).  -> weights [B, heads] fp32, 1/k on the k selected heads."""
    w = np.zeros((batch, num_heads), dtype=np.float32)
    for i in range(batch):
        num_rows = rng.randint(1, 7)
        sel = rng.choice(num_heads, size=num_rows, replace=False)
        w[i, sel] = 1.0 / num_rows
    return torch.from_numpy(w)


@torch.no_grad()
def pseudo_masks(encoder, x: torch.Tensor, rand: bool = False, head_weights: torch.Tensor | None = None, which: int = 0) -> torch.Tensor:
    """y [B, 1, S, S] fp32 in {0, 1}: ``output / 255`` of utils.threshold for every image of x [B, C, S, S] (SSS/PGT.py:55-91).
    rand: average a random subset of heads per image (``select_heads``); head_weights [B, heads] overrides the draw.
    which: 0 = "ours" (the mask PGT trains on), 1 = Otsu of the image, 2 = Otsu of the attention map."""
    rows = encoder.cls_attention_rows(x)                       # [B, heads, N]
    B, _, S, S2 = x.shape
    if S != S2:
        raise ValueError("pseudo_masks expects square tiles")
    p = encoder.patch_embed.patch_size
    lh = S // p
    if rand and head_weights is None:
        head_weights = select_heads(B, rows.shape[1])
    if head_weights is None:
        low = head_mean_maps(rows, per_tile_minmax255=False)   # np.mean over all heads (SSS/PGT.py:75)
    else:
        w = head_weights.to(device=rows.device, dtype=torch.float32)
        low = (rows[:, :, 1:] * w[:, :, None]).sum(1).contiguous()
    masks = torch.empty(B, 3, S, S, dtype=torch.uint8, device=x.device)
    thr = torch.empty(B, 3, dtype=torch.int32, device=x.device)
    xx = x.detach().to(torch.float32).contiguous()
    check(_lib.load_library().vitocm_tile_threshold(ptr(low), ptr(xx), B, x.shape[1], S, lh, lh, ptr(masks), ptr(thr), None, None, None,
                                                    cur_stream()))
    return masks[:, which:which + 1].to(torch.float32).div_(255.0)
