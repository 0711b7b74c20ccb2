"""Helpers for the GPU tests: thin wrappers over the kernel-level C-ABI entry points."""
import ctypes as C

import torch

import vitocm_b200 as vob
from vitocm_b200._lib import VitocmConfig, check, cur_stream, ptr


def make_engine(embed_dim=128, heads=2, depth=1, hidden=512, patch=8, chans=3, precision=0):
    lib = vob._lib.load_library()
    cfg = VitocmConfig(embed_dim, depth, heads, hidden, patch, chans, 1e-6, 0.125, precision)
    h = C.c_void_p()
    check(lib.vitocm_create(C.byref(cfg), C.byref(h)))
    return h


def split_bf16(x: torch.Tensor) -> torch.Tensor:
    """fp32 [R, K] -> bf16 [R, 2K] = hi | lo"""
    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16)
    return torch.cat([hi, lo], dim=1).contiguous()


def gemm(engine, A, B, M, N, K, split_in, epi, bias, out, ldo, split_out=0, lo_off=0):
    lib = vob._lib.load_library()
    check(lib.vitocm_gemm(engine, ptr(A), A.stride(0), ptr(B), B.stride(0), M, N, K, split_in, epi, ptr(bias), ptr(out), ldo,
                          split_out, lo_off, cur_stream()))
    torch.cuda.synchronize()


def attention(engine, qkv, B, N, ctx):
    lib = vob._lib.load_library()
    check(lib.vitocm_attention(engine, ptr(qkv), qkv.stride(0), B, N, ptr(ctx), ctx.stride(0), cur_stream()))
    torch.cuda.synchronize()


def attention_reference(q, k, v, scale):
    """q, k, v: [B, H, N, 64] fp32 -> ctx [B, N, H*64] fp32 (SSS/dino/vision_transformer.py:83-87)."""
    a = (q @ k.transpose(-2, -1)) * scale
    a = a.softmax(dim=-1)
    o = a @ v
    B, H, N, dh = o.shape
    return o.transpose(1, 2).reshape(B, N, H * dh)


def build_model(cfg, sd, precision, chunk_tiles=16):
    """vitocm VisionTransformer for an oracle ViTConfig + state dict."""
    from functools import partial
    m = vob.VisionTransformer(img_size=[cfg.img_size], patch_size=cfg.patch_size, in_chans=cfg.in_chans, num_classes=0,
                              embed_dim=cfg.embed_dim, depth=cfg.depth, num_heads=cfg.num_heads, mlp_ratio=cfg.mlp_ratio,
                              qkv_bias=True, norm_layer=partial(torch.nn.LayerNorm, eps=cfg.eps), precision=precision,
                              chunk_tiles=chunk_tiles)
    m.load_state_dict(sd, strict=True)
    return m.cuda().eval()
