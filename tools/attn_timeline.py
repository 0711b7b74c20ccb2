"""Per-kv-block timeline of the attention kernel (SM clocks) for two co-scheduled CTAs, at full-chip load."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import vitocm_b200 as vob
from vitocm_b200._lib import check, cur_stream, ptr
from gpu_util import make_engine

B, H, N = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (32, 6, 785)
D = 64 * H
eng = make_engine(embed_dim=D, heads=H, precision=0)
qkv = (torch.randn(B * N, 3 * D, device="cuda") * 1.0).to(torch.bfloat16)
ctx = torch.empty(B * N, D, device="cuda", dtype=torch.bfloat16)
stamps = torch.zeros(2, 2, 16, 8, dtype=torch.int64, device="cuda")
lib = vob._lib.load_library()
for _ in range(3):
    check(lib.vitocm_attention_timeline(eng, ptr(qkv), qkv.stride(0), B, N, ptr(ctx), ctx.stride(0), ptr(stamps), cur_stream()))
torch.cuda.synchronize()
s = stamps.cpu()
t0 = int(s[s > 0].min())
names = ["wait S", "S done", "S in regs", "max done", "PV(j-1) done", "exps issued", "P handed"]
for cta in range(2):
    print(f"== CTA (q tile {cta}) softmax warp 0, clocks relative to first stamp")
    for j in range(7):
        ev = [int(s[cta, 0, j, k]) - t0 if s[cta, 0, j, k] > 0 else -1 for k in range(7)]
        mma = [int(s[cta, 1, j, k]) - t0 if s[cta, 1, j, k] > 0 else -1 for k in range(2)]
        print(f" j={j} " + " ".join(f"{n}={v}" for n, v in zip(names, ev)) + f" | MMA: S issued={mma[0]} PV issued={mma[1]}")

for cta in range(2):
    a, b = int(s[cta, 0, 15, 6]), int(s[cta, 0, 15, 7])
    items = (B * H * ((N + 127) // 128) - cta + 295) // 296
    print(f"CTA {cta}: lifetime {b - a} clk for {items} work items = {(b - a) / max(items, 1):.0f} clk per item")
