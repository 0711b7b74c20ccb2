"""Microbenchmark: attention forward (+LSE) and backward at the MIM training shape; per-class device times."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import vitocm_b200 as vob
from gpu_util import make_engine
from vitocm_b200._lib import check, cur_stream, ptr

B, H, N = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (64, 6, 785)
lib = vob._lib.load_library()
eng = make_engine(embed_dim=64 * H, heads=H, precision=0)
D = 64 * H
g = torch.Generator(device="cuda").manual_seed(0)
qkv = torch.randn(B * N, 3 * D, generator=g, device="cuda").to(torch.bfloat16)
dctx = (torch.randn(B * N, D, generator=g, device="cuda") * 0.5).to(torch.bfloat16)
ctx = torch.empty(B * N, D, device="cuda", dtype=torch.bfloat16)
lse = torch.empty(B, H, (N + 127) // 128 * 128, device="cuda")
dqkv = torch.empty(B * N, 3 * D, device="cuda", dtype=torch.bfloat16)
delta = torch.empty(B, H, (N + 127) // 128 * 128, device="cuda")
dqacc = torch.zeros(B * N, D, device="cuda")
def fwd():
    check(lib.vitocm_attention_fwd_lse(eng, ptr(qkv), qkv.stride(0), B, N, ptr(ctx), ctx.stride(0), ptr(lse), cur_stream()))
def bwd():
    check(lib.vitocm_attention_bwd(eng, ptr(qkv), qkv.stride(0), ptr(ctx), ptr(dctx), dctx.stride(0), ptr(lse), ptr(delta), ptr(dqacc),
                                   ptr(dqkv), dqkv.stride(0), B, N, cur_stream()))
for _ in range(3):
    fwd(); bwd()
torch.cuda.synchronize()
vob._lib.profile_enable(True)
for _ in range(10):
    fwd(); bwd()
torch.cuda.synchronize()
prof = vob._lib.profile_read()
flops_f = 4.0 * N * N * D * B
for k, (ms, n) in prof.items():
    if n:
        fl = flops_f if k == "attention" else 2.5 * flops_f if k == "attention_bwd" else 0
        print(f"{k:20s} {ms / n * 1e3:9.1f} us/launch x{n}" + (f"  {fl / (ms / n) / 1e9:7.1f} TFLOP/s" if fl else ""), flush=True)
if os.environ.get("VITOCM_ABW_DEBUG") == "2":
    import numpy as np
    bwd(); torch.cuda.synchronize()
    tl = np.zeros(2 * 8 * 8, dtype=np.int64)
    check(lib.vitocm_debug_abw_timeline(tl.ctypes.data))
    tl = tl.reshape(2, 8, 8)
    t0 = tl[0, 0, 0]
    print("softmax warp 0: per query tile [wait S/dP, S/dP ready, P/dS computed, dQ(i-1) drained, P/dS handed over | dq_full(i-1) seen]")
    for i in range(7):
        print("  i=%d " % i, [int(v - t0) if v else None for v in tl[0, i, :7]])
    print("MMA thread: [waiting for P/dS, P/dS ready, all MMAs of tile issued]")
    for i in range(7):
        print("  i=%d " % i, [int(v - t0) if v else None for v in tl[1, i, :3]])
