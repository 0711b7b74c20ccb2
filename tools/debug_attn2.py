import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import vitocm_b200 as vob
from gpu_util import attention, attention_reference, make_engine, split_bf16
B, H, N = 1, 1, 64
D = 64
for flag in (0, 1, 2, 4, 6, 7, 8, 15, 16, 32, 48):
    os.environ["VITOCM_ATTN_DEBUG"] = str(flag)
    eng = make_engine(embed_dim=64, heads=1, precision=1)
    g = torch.Generator(device="cuda").manual_seed(1)
    q, k, v = (torch.randn(B, H, N, 64, generator=g, device="cuda") for _ in range(3))
    qkv32 = torch.stack([q, k, v], 0).permute(1, 3, 0, 2, 4).reshape(B * N, 3 * D).contiguous()
    qkv = split_bf16(qkv32); src = qkv[:, :3 * D].float() + qkv[:, 3 * D:].float()
    s5 = src.reshape(B, N, 3, H, 64).permute(2, 0, 3, 1, 4)
    ref = attention_reference(s5[0], s5[1], s5[2], 0.125).reshape(B * N, D)
    ctx = torch.full((B * N, 2 * D), 7.0, device="cuda", dtype=torch.bfloat16)
    attention(eng, qkv, B, N, ctx)
    got = ctx[:, :D].float() + ctx[:, D:].float()
    nan = torch.isnan(got)
    print(f"flag={flag:2d} nan={nan.sum().item():5d} row0: {got[0,:4].tolist()} ref {ref[0,:4].tolist()}")
