"""Drop-in for the MIM half of the reference's ``model.py`` on B200 (forward / loss evaluation).

Mirrors ``VisionTransformerForSimMIM`` (SSS/model.py:11-53), ``MIM`` (:55-89) and ``build_model``
(:91-108): same constructor arguments, parameter names (``mask_token``, ``decoder.0.{weight,bias}``) and
return values.  The arithmetic runs in libvitocm.so (``vitocm_mim_forward``): patch-embedding GEMM with the
mask-token mix fused in its epilogue, the transformer blocks, final norm, the 1x1-conv decoder as a
per-token tcgen05 GEMM, PixelShuffle + masked L1 in one pass.  There is no autograd here: the backward
kernels of the training step (SSS/mim.py:153-180) are not built yet, so ``loss`` carries no graph.
``MaskGenerator`` (SSS/data.py:163-186) is host-side numpy in the reference and stays so.
"""
from __future__ import annotations

import math
from functools import partial

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from ._lib import check, cur_stream, ptr
from .vision_transformer import VisionTransformer


class MaskGenerator:
    """SSS/data.py:163-186 (uses numpy's global RNG, like the reference)."""

    def __init__(self, input_size=192, mask_patch_size=32, model_patch_size=4, mask_ratio=0.6):
        assert input_size % mask_patch_size == 0 and mask_patch_size % model_patch_size == 0
        self.input_size, self.mask_patch_size, self.model_patch_size, self.mask_ratio = input_size, mask_patch_size, model_patch_size, mask_ratio
        self.rand_size = input_size // mask_patch_size
        self.scale = mask_patch_size // model_patch_size
        self.token_count = self.rand_size ** 2
        self.mask_count = int(np.ceil(self.token_count * mask_ratio))

    def __call__(self):
        idx = np.random.permutation(self.token_count)[:self.mask_count]
        mask = np.zeros(self.token_count, dtype=int)
        mask[idx] = 1
        return mask.reshape(self.rand_size, self.rand_size).repeat(self.scale, axis=0).repeat(self.scale, axis=1)


class VisionTransformerForSimMIM(VisionTransformer):
    """SSS/model.py:11-53.  As in the reference, ``img_size`` is *not* forwarded to the base constructor (the
    position table always has 28*28+1 rows for patch 8) and is bicubically resized whenever img_size[0] != 224."""

    def __init__(self, interpolate_encoding=False, img_size=224, **kwargs):
        super().__init__(**kwargs)
        self.mask_token = nn.Parameter(torch.zeros(1, 1, self.embed_dim))
        self.img_size = img_size
        self._trunc_normal_(self.mask_token, std=.02)
        self.interpolate_encoding = interpolate_encoding

    def _trunc_normal_(self, tensor, mean=0., std=1.):
        cdf = lambda v: (1.0 + math.erf(v / math.sqrt(2.0))) / 2.0   # noqa: E731  (a = -std, b = std)
        lo, hi = cdf(-1.0), cdf(1.0)
        with torch.no_grad():
            tensor.uniform_(2 * lo - 1, 2 * hi - 1).erfinv_().mul_(std * math.sqrt(2.0)).add_(mean).clamp_(min=-std, max=std)
        return tensor

    def _mim_pos(self, x):
        size = self.img_size[0] if isinstance(self.img_size, (list, tuple)) else self.img_size
        B, _, H, W = x.shape
        n = (H // self.patch_embed.patch_size) * (W // self.patch_embed.patch_size)
        if size != 224:
            return self._pos_table(n, size, size)
        return self._pos_table(self.pos_embed.shape[1] - 1, 224, 224)

    @torch.no_grad()
    def forward(self, x, mask):
        """-> [B, D, H/p, W/p] (model.py:25-53)."""
        assert mask is not None
        xx = self._check_input(x)
        eng = self._ensure_engine()
        B, _, H, W = xx.shape
        N = self._tokens(xx)
        pos = self._mim_pos(xx)
        assert pos.shape[0] == N, "position table does not match the token count (img_size vs input size)"
        X = torch.empty(B, N, self.embed_dim, dtype=torch.float32, device=xx.device)
        m = mask.detach().reshape(B, -1).to(device=xx.device, dtype=torch.float32).contiguous()
        check(_lib.load_library().vitocm_prepare_tokens(eng, ptr(xx), B, H, W, ptr(pos), ptr(m), ptr(X), cur_stream()))
        self._run_blocks(X, 0, self.depth)
        z = self._final_norm(X)[:, 1:]
        h = w = int((N - 1) ** 0.5)
        return z.permute(0, 2, 1).reshape(B, self.embed_dim, h, w)     # a view change, no arithmetic

    def no_weight_decay(self):
        return {"pos_embed", "cls_token", "mask_token"}


class MIM(nn.Module):
    """SSS/model.py:55-89: encoder + (1x1 conv, PixelShuffle) decoder + masked L1 loss."""

    def __init__(self, encoder, encoder_stride):
        super().__init__()
        self.encoder = encoder
        self.encoder_stride = encoder_stride
        self.decoder = nn.Sequential(
            nn.Conv2d(in_channels=self.encoder.num_features, out_channels=self.encoder_stride ** 2 * 3, kernel_size=1),
            nn.PixelShuffle(self.encoder_stride),
        )
        self.in_chans = 3
        self.patch_size = 8
        if encoder_stride != encoder.patch_embed.patch_size:
            raise NotImplementedError("vitocm MIM: encoder_stride must equal the patch size")
        # the decoder's parameters are loaded into the encoder's engine under their state-dict names
        self.encoder._extra_engine_params = lambda: [("decoder.0.weight", self.decoder[0].weight), ("decoder.0.bias", self.decoder[0].bias)]

    @torch.no_grad()
    def forward(self, x, mask):
        """-> (loss, x_rec, mask upsampled to pixels) (model.py:71-77)."""
        enc = self.encoder
        xx = enc._check_input(x)
        eng = enc._ensure_engine()
        B, C, H, W = xx.shape
        N = enc._tokens(xx)
        pos = enc._mim_pos(xx)
        assert pos.shape[0] == N, "position table does not match the token count (img_size vs input size)"
        m = mask.detach().reshape(B, -1).to(device=xx.device, dtype=torch.float32).contiguous()
        x_rec = torch.empty_like(xx)
        sums = torch.empty(2, dtype=torch.float64, device=xx.device)
        chunk = max(1, min(enc.chunk_tiles, B))
        ws = enc._workspace(chunk, N, xx.device)
        check(_lib.load_library().vitocm_mim_forward(eng, ptr(xx), B, H, W, ptr(pos), ptr(m), ptr(x_rec), ptr(sums), ptr(ws),
                                                     ws.numel(), chunk, cur_stream()))
        loss = (sums[0] / (sums[1] + 1e-5) / self.in_chans).to(torch.float32)
        mask_up = mask.repeat_interleave(self.patch_size, 1).repeat_interleave(self.patch_size, 2).unsqueeze(1).contiguous()
        return loss, x_rec, mask_up

    @torch.jit.ignore
    def no_weight_decay(self):
        if hasattr(self.encoder, 'no_weight_decay'):
            return {'encoder.' + i for i in self.encoder.no_weight_decay()}
        return {}

    @torch.jit.ignore
    def no_weight_decay_keywords(self):
        if hasattr(self.encoder, 'no_weight_decay_keywords'):
            return {'encoder.' + i for i in self.encoder.no_weight_decay_keywords()}
        return {}


def build_model(args, depth=12, num_heads=6, precision="bf16"):
    """SSS/model.py:91-108.  The reference currently hard-codes an experimental depth=4 / num_heads=3 (head_dim 128);
    the shipped training log (SSS/output/log_rank0.txt) is the 12-block, 6-head ViT-S/8 used here by default --
    the kernels are specialised for head_dim 64."""
    return VisionTransformerForSimMIM(patch_size=args.MODEL.PATCH_SIZE, embed_dim=384, depth=depth, num_heads=num_heads,
                                      mlp_ratio=4, img_size=[args.DATA.IMG_SIZE], qkv_bias=True,
                                      norm_layer=partial(nn.LayerNorm, eps=1e-6), interpolate_encoding=True,
                                      precision=precision)
