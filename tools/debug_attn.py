import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import vitocm_b200 as vob
from gpu_util import attention, attention_reference, make_engine, split_bf16

for precision in (1, 0):
    for (B, H, N) in ((1, 1, 48), (1, 1, 64), (1, 1, 65), (1, 1, 80), (1, 1, 96), (1, 1, 128), (1, 1, 256)):
        eng = make_engine(embed_dim=64 * H, heads=H, precision=precision)
        D = 64 * H
        g = torch.Generator(device="cuda").manual_seed(1)
        q, k, v = (torch.randn(B, H, N, 64, generator=g, device="cuda") for _ in range(3))
        qkv32 = torch.stack([q, k, v], 0).permute(1, 3, 0, 2, 4).reshape(B * N, 3 * D).contiguous()
        if precision:
            qkv = split_bf16(qkv32); src = qkv[:, :3 * D].float() + qkv[:, 3 * D:].float()
        else:
            qkv = qkv32.to(torch.bfloat16).contiguous(); src = qkv.float()
        s5 = src.reshape(B, N, 3, H, 64).permute(2, 0, 3, 1, 4)
        ref = attention_reference(s5[0], s5[1], s5[2], 0.125).reshape(B * N, D)
        parts = 2 if precision else 1
        ctx = torch.full((B * N, D * parts), 7.0, device="cuda", dtype=torch.bfloat16)
        try:
            attention(eng, qkv, B, N, ctx)
        except Exception as e:
            print("EXC", precision, N, e); sys.exit(0)
        got = ctx[:, :D].float() + (ctx[:, D:].float() if precision else 0)
        nan = torch.isnan(got)
        ok = (~nan).sum().item()
        print(f"prec={precision} N={N} nan rows={nan.any(1).sum().item()} nan cols={nan.any(0).sum().item()} err(non-nan)={(got-ref)[~nan].abs().max().item() if ok else -1:.3e}")
        print("   sample hi:", ctx[0, :4].float().tolist(), " lo:", ctx[0, D:D+4].float().tolist() if precision else None, " ref:", ref[0, :4].tolist())
        if nan.any():
            print("  nan row idx:", nan.any(1).nonzero().flatten()[:20].tolist(), " nan col idx:", nan.any(0).nonzero().flatten()[:20].tolist())
            print("  hi part nan:", torch.isnan(ctx[:, :D].float()).sum().item(), " lo part nan:", torch.isnan(ctx[:, D:].float()).sum().item() if precision else 0)
