"""Per-kv-block timeline (SM clocks) of pipelines 0 and 1 of CTA 0 of the four-pipeline attention kernel, at full-chip load.
usage: python tools/attn_quad_timeline.py [B H N]   (env PRECISION 0 | 2, VITOCM_ATTN_TL_ITEM)"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import vitocm_b200 as vob
from vitocm_b200._lib import check, cur_stream, ptr
from gpu_util import make_engine

B, H, N = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (175, 6, 785)
D = 64 * H
PREC = int(os.environ.get("PRECISION", "2"))
DT = torch.float16 if PREC == 2 else torch.bfloat16
eng = make_engine(embed_dim=D, heads=H, precision=PREC)
qkv = torch.randn(B * N, 3 * D, device="cuda").to(DT)
ctx = torch.empty(B * N, D, device="cuda", dtype=DT)
stamps = torch.zeros(2, 2, 16, 8, dtype=torch.int64, device="cuda")
lib = vob._lib.load_library()
for _ in range(3):
    check(lib.vitocm_attention_timeline(eng, ptr(qkv), qkv.stride(0), B, N, ptr(ctx), ctx.stride(0), ptr(stamps), cur_stream()))
torch.cuda.synchronize()
s = stamps.cpu()
t0 = int(s[s > 0].min())
names = ["wait S", "S done", "c0 regs", "c0 exps", "c1 regs", "c1 exps", "P handed"]
nkv = (N + 63) // 64
for pipe in range(2):
    print(f"== pipeline {pipe} of CTA 0, softmax warp 0; clocks relative to the first stamp")
    for j in range(min(nkv, 16)):
        ev = [int(s[pipe, 0, j, k]) - t0 if s[pipe, 0, j, k] > 0 else -1 for k in range(7)]
        mma = [int(s[pipe, 1, j, k]) - t0 if s[pipe, 1, j, k] > 0 else -1 for k in range(3)]
        print(f" j={j:2d} " + " ".join(f"{n}={v}" for n, v in zip(names, ev)) + f" | ctrl: S issued={mma[0]} P seen={mma[1]} PV issued={mma[2]}")
