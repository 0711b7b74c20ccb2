"""Oracle (test infrastructure, see oracle/__init__.py): torch-CPU fp32 restatement of the
reference DINO ViT forward and of the SimMIM wrapper, written functionally over a
state-dict so that it does not share code with the product's nn.Module.

Reference: /root/reference/Self-supervised_segmentation (abbrev. SSS)
  SSS/dino/vision_transformer.py  (ViT, attention getters)
  SSS/model.py                    (VisionTransformerForSimMIM, MIM)
  SSS/dino/utils.py:482-520       (trunc_normal_)
  SSS/data.py:163-186             (MaskGenerator)
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch
import torch.nn.functional as F


@dataclass(frozen=True)
class ViTConfig:
    """Constructor constants of SSS/dino/vision_transformer.py:137-139 and the factories :259-279."""
    embed_dim: int = 384
    depth: int = 12
    num_heads: int = 6
    mlp_ratio: float = 4.0
    patch_size: int = 8
    in_chans: int = 3
    img_size: int = 224
    eps: float = 1e-6          # partial(nn.LayerNorm, eps=1e-6), vit.py:262/269/278

    @property
    def head_dim(self) -> int:
        return self.embed_dim // self.num_heads

    @property
    def hidden(self) -> int:
        return int(self.embed_dim * self.mlp_ratio)


VIT_TINY = dict(embed_dim=192, depth=12, num_heads=3)     # vit.py:259-263
VIT_SMALL = dict(embed_dim=384, depth=12, num_heads=6)    # vit.py:266-270
VIT_BASE = dict(embed_dim=768, depth=12, num_heads=12)    # vit.py:275-279


# --------------------------------------------------------------------------------------
# initialisation (so that oracle-side weights can be produced without the reference)
# --------------------------------------------------------------------------------------
def _trunc_normal_(t: torch.Tensor, std: float, a: float = -2.0, b: float = 2.0) -> torch.Tensor:
    """SSS/dino/utils.py:482-515 (mean 0): uniform -> erfinv -> scale -> clamp."""
    def cdf(x):
        return (1.0 + math.erf(x / math.sqrt(2.0))) / 2.0
    lo, hi = cdf(a / std), cdf(b / std)
    with torch.no_grad():
        t.uniform_(2 * lo - 1, 2 * hi - 1)
        t.erfinv_()
        t.mul_(std * math.sqrt(2.0))
        t.clamp_(min=a, max=b)
    return t


def init_state_dict(cfg: ViTConfig, seed: int = 0, mim: bool = False) -> dict[str, torch.Tensor]:
    """Random-init weights with the reference's distributions (vit.py:152-174): trunc-normal
    std .02 for Linear weights / pos_embed / cls_token, zero biases, LayerNorm (1, 0); the
    patch-embed conv keeps torch's default (kaiming-uniform) init.  The draw ORDER is the
    oracle's own -- tests never rely on it matching the reference's RNG stream; golden
    fixtures carry the weights (tiny model) or weight checksums (ViT-S)."""
    g = torch.Generator().manual_seed(seed)
    D, p, C = cfg.embed_dim, cfg.patch_size, cfg.in_chans
    n = (cfg.img_size // p) ** 2
    sd: dict[str, torch.Tensor] = {}

    def tn(*shape):
        t = torch.empty(*shape)
        def cdf(x):
            return (1.0 + math.erf(x / math.sqrt(2.0))) / 2.0
        std = 0.02
        lo, hi = cdf(-2.0 / std), cdf(2.0 / std)
        t.uniform_(2 * lo - 1, 2 * hi - 1, generator=g)
        t.erfinv_().mul_(std * math.sqrt(2.0)).clamp_(-2.0, 2.0)
        return t

    fan_in = C * p * p
    bound = 1.0 / math.sqrt(fan_in)
    sd["cls_token"] = tn(1, 1, D)
    sd["pos_embed"] = tn(1, n + 1, D)
    sd["patch_embed.proj.weight"] = (torch.rand(D, C, p, p, generator=g) * 2 - 1) * bound
    sd["patch_embed.proj.bias"] = (torch.rand(D, generator=g) * 2 - 1) * bound
    for i in range(cfg.depth):
        pre = f"blocks.{i}."
        sd[pre + "norm1.weight"] = torch.ones(D)
        sd[pre + "norm1.bias"] = torch.zeros(D)
        sd[pre + "attn.qkv.weight"] = tn(3 * D, D)
        sd[pre + "attn.qkv.bias"] = torch.zeros(3 * D)
        sd[pre + "attn.proj.weight"] = tn(D, D)
        sd[pre + "attn.proj.bias"] = torch.zeros(D)
        sd[pre + "norm2.weight"] = torch.ones(D)
        sd[pre + "norm2.bias"] = torch.zeros(D)
        sd[pre + "mlp.fc1.weight"] = tn(cfg.hidden, D)
        sd[pre + "mlp.fc1.bias"] = torch.zeros(cfg.hidden)
        sd[pre + "mlp.fc2.weight"] = tn(D, cfg.hidden)
        sd[pre + "mlp.fc2.bias"] = torch.zeros(D)
    sd["norm.weight"] = torch.ones(D)
    sd["norm.bias"] = torch.zeros(D)
    if mim:
        t = torch.empty(1, 1, D)
        # model.py:22-23: trunc_normal_(mask_token, std=.02, a=-std, b=std)
        def cdf(x):
            return (1.0 + math.erf(x / math.sqrt(2.0))) / 2.0
        lo, hi = cdf(-1.0), cdf(1.0)
        t.uniform_(2 * lo - 1, 2 * hi - 1, generator=g)
        t.erfinv_().mul_(0.02 * math.sqrt(2.0)).clamp_(-0.02, 0.02)
        sd["mask_token"] = t
    return sd


def randomize_affine(sd: dict[str, torch.Tensor], seed: int = 1, scale: float = 0.1) -> dict[str, torch.Tensor]:
    """Perturb biases and LayerNorm affines so tests exercise them (random init leaves them 0/1)."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k, v in sd.items():
        if k.endswith(".bias") and "patch_embed" not in k:
            out[k] = v + scale * torch.randn(v.shape, generator=g)
        elif "norm" in k and k.endswith(".weight"):
            out[k] = v + scale * torch.randn(v.shape, generator=g)
        else:
            out[k] = v.clone()
    return out


def param_count(sd: dict[str, torch.Tensor]) -> int:
    return int(sum(v.numel() for v in sd.values()))


# --------------------------------------------------------------------------------------
# forward pieces
# --------------------------------------------------------------------------------------
def interpolate_pos_encoding(sd, cfg: ViTConfig, npatch: int, w: int, h: int) -> torch.Tensor:
    """vit.py:176-196.  Bicubic resize of the patch position table with the +0.1 trick."""
    pos = sd["pos_embed"]
    N = pos.shape[1] - 1
    if npatch == N and w == h:
        return pos
    class_pos = pos[:, 0]
    patch_pos = pos[:, 1:]
    dim = pos.shape[-1]
    w0 = w // cfg.patch_size + 0.1
    h0 = h // cfg.patch_size + 0.1
    s = int(math.sqrt(N))
    patch_pos = F.interpolate(patch_pos.reshape(1, s, s, dim).permute(0, 3, 1, 2),
                              scale_factor=(w0 / math.sqrt(N), h0 / math.sqrt(N)), mode="bicubic")
    assert int(w0) == patch_pos.shape[-2] and int(h0) == patch_pos.shape[-1]
    patch_pos = patch_pos.permute(0, 2, 3, 1).reshape(1, -1, dim)
    return torch.cat((class_pos.unsqueeze(0), patch_pos), dim=1)


def patch_embed(sd, cfg: ViTConfig, x: torch.Tensor) -> torch.Tensor:
    """vit.py:129-132: conv k=p, s=p (+bias) -> flatten(2).transpose(1,2) -> [B, n, D]."""
    y = F.conv2d(x, sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"], stride=cfg.patch_size)
    return y.flatten(2).transpose(1, 2)


def prepare_tokens(sd, cfg: ViTConfig, x: torch.Tensor) -> torch.Tensor:
    """vit.py:198-209."""
    B, _, w, h = x.shape
    t = patch_embed(sd, cfg, x)
    t = torch.cat((sd["cls_token"].expand(B, -1, -1), t), dim=1)
    return t + interpolate_pos_encoding(sd, cfg, t.shape[1] - 1, w, h)


def _ln(x, w, b, eps):
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


def attention(sd, cfg: ViTConfig, i: int, x: torch.Tensor):
    """vit.py:78-90 -> (y, attn[B,H,N,N], qkv[3,B,H,N,dh])."""
    pre = f"blocks.{i}.attn."
    B, N, C = x.shape
    H = cfg.num_heads
    qkv = F.linear(x, sd[pre + "qkv.weight"], sd[pre + "qkv.bias"]).reshape(B, N, 3, H, C // H).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    attn = (q @ k.transpose(-2, -1)) * (cfg.head_dim ** -0.5)
    attn = attn.softmax(dim=-1)
    y = (attn @ v).transpose(1, 2).reshape(B, N, C)
    y = F.linear(y, sd[pre + "proj.weight"], sd[pre + "proj.bias"])
    return y, attn, qkv


def mlp(sd, cfg: ViTConfig, i: int, x: torch.Tensor) -> torch.Tensor:
    """vit.py:57-63: fc1 -> exact-erf GELU -> fc2."""
    pre = f"blocks.{i}.mlp."
    h = F.gelu(F.linear(x, sd[pre + "fc1.weight"], sd[pre + "fc1.bias"]))
    return F.linear(h, sd[pre + "fc2.weight"], sd[pre + "fc2.bias"])


def block(sd, cfg: ViTConfig, i: int, x: torch.Tensor):
    """vit.py:106-114 -> (x_out, attn, qkv)."""
    pre = f"blocks.{i}."
    y, attn, qkv = attention(sd, cfg, i, _ln(x, sd[pre + "norm1.weight"], sd[pre + "norm1.bias"], cfg.eps))
    x = x + y
    x = x + mlp(sd, cfg, i, _ln(x, sd[pre + "norm2.weight"], sd[pre + "norm2.bias"], cfg.eps))
    return x, attn, qkv


@torch.no_grad()
def get_last_selfattention(sd, cfg: ViTConfig, x: torch.Tensor) -> torch.Tensor:
    """vit.py:239-246 -> attn of the last block, [B,H,N,N]."""
    t = prepare_tokens(sd, cfg, x)
    for i in range(cfg.depth - 1):
        t, _, _ = block(sd, cfg, i, t)
    pre = f"blocks.{cfg.depth - 1}."
    _, attn, _ = attention(sd, cfg, cfg.depth - 1, _ln(t, sd[pre + "norm1.weight"], sd[pre + "norm1.bias"], cfg.eps))
    return attn


@torch.no_grad()
def get_intermediate_feat(sd, cfg: ViTConfig, x: torch.Tensor, n: int = 1):
    """vit.py:225-237 -> (feat list, attn list, qkv list) for the last n blocks."""
    t = prepare_tokens(sd, cfg, x)
    feat, attns, qkvs = [], [], []
    for i in range(cfg.depth):
        t, attn, qkv = block(sd, cfg, i, t)
        if cfg.depth - i <= n:
            feat.append(_ln(t, sd["norm.weight"], sd["norm.bias"], cfg.eps))
            qkvs.append(qkv)
            attns.append(attn)
    return feat, attns, qkvs


@torch.no_grad()
def forward_feats(sd, cfg: ViTConfig, x: torch.Tensor) -> torch.Tensor:
    """vit.py:218-223 -> norm(x) [B,N,D]; forward() (:211-216) is [:, 0] of this."""
    t = prepare_tokens(sd, cfg, x)
    for i in range(cfg.depth):
        t, _, _ = block(sd, cfg, i, t)
    return _ln(t, sd["norm.weight"], sd["norm.bias"], cfg.eps)


@torch.no_grad()
def cls_attention_rows(sd, cfg: ViTConfig, x: torch.Tensor) -> torch.Tensor:
    """The only slice the hot-path callers read: attn[:, :, 0, :] -> [B,H,N]
    (SSS/utils.py:232 with query=0, SSS/eval.py:137, SSS/sw_processing.py:240)."""
    return get_last_selfattention(sd, cfg, x)[:, :, 0, :].contiguous()


# --------------------------------------------------------------------------------------
# SimMIM wrapper (model.py)
# --------------------------------------------------------------------------------------
def mask_generator(rng: np.random.RandomState, input_size=224, mask_patch_size=16, model_patch_size=8,
                   mask_ratio=0.5) -> np.ndarray:
    """SSS/data.py:163-186 (uses the numpy global RNG there; a RandomState here)."""
    rand_size = input_size // mask_patch_size
    scale = mask_patch_size // model_patch_size
    count = rand_size ** 2
    mask_count = int(np.ceil(count * mask_ratio))
    idx = rng.permutation(count)[:mask_count]
    mask = np.zeros(count, dtype=int)
    mask[idx] = 1
    mask = mask.reshape(rand_size, rand_size)
    return mask.repeat(scale, axis=0).repeat(scale, axis=1)


def simmim_encoder(sd, cfg: ViTConfig, x: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """model.py:25-53 -> [B, D, h, w] (autograd-capable: no no_grad here)."""
    t = patch_embed(sd, cfg, x)
    B, L, _ = t.shape
    w = mask.flatten(1).unsqueeze(-1).type_as(t)
    t = t * (1 - w) + sd["mask_token"].expand(B, L, -1) * w
    t = torch.cat((sd["cls_token"].expand(B, -1, -1), t), dim=1)
    if cfg.img_size != 224:
        t = t + interpolate_pos_encoding(sd, cfg, L, cfg.img_size, cfg.img_size)
    else:
        t = t + sd["pos_embed"]
    for i in range(cfg.depth):
        t, _, _ = block(sd, cfg, i, t)
    t = _ln(t, sd["norm.weight"], sd["norm.bias"], cfg.eps)[:, 1:]
    Hh = int(L ** 0.5)
    return t.permute(0, 2, 1).reshape(B, -1, Hh, Hh)


def mim_forward(sd, cfg: ViTConfig, dec_w: torch.Tensor, dec_b: torch.Tensor, x: torch.Tensor, mask: torch.Tensor):
    """model.py:71-77 -> (loss, x_rec, mask_up).  Decoder = 1x1 conv D->p*p*3 + PixelShuffle(p) (:61-66)."""
    z = simmim_encoder(sd, cfg, x, mask)
    x_rec = F.pixel_shuffle(F.conv2d(z, dec_w, dec_b), cfg.patch_size)
    m = mask.repeat_interleave(cfg.patch_size, 1).repeat_interleave(cfg.patch_size, 2).unsqueeze(1).contiguous()
    loss_recon = F.l1_loss(x, x_rec, reduction="none")
    loss = (loss_recon * m).sum() / (m.sum() + 1e-5) / cfg.in_chans
    return loss, x_rec, m


# --------------------------------------------------------------------------------------
# synthetic inputs: dark, right-skewed, spatially correlated gray field
# --------------------------------------------------------------------------------------
def synthetic_gray(size: int, seed: int = 1234, batch: int = 1) -> torch.Tensor:
    """[B,1,S,S] in [0,1] quantised to k/255 (the reference feeds ToTensor() of u8 images)."""
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
    """[B,3,S,S] fp32, R=G=B (real OCM images are gray)."""
    return synthetic_gray(size, seed, batch).expand(-1, 3, -1, -1).contiguous()


def synthetic_mosaic_u8(size: int, seed: int = 4321) -> np.ndarray:
    """[S,S] uint8 gray mosaic."""
    g = synthetic_gray(size, seed, 1)[0, 0]
    return (g * 255.0).round().to(torch.uint8).numpy()
