"""Diagnostics: attention backward per-part errors for a few shapes (each in its own process via argv)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import vitocm_b200 as vob
from gpu_util import make_engine
from vitocm_b200._lib import check, cur_stream, ptr
from test_gpu_train_kernels import _attention_autograd, _rand

B, H, N = (int(v) for v in sys.argv[1:4])
lib = vob._lib.load_library()
eng = make_engine(embed_dim=64 * H, heads=H, precision=0)
D = 64 * H
qkv = _rand((B * N, 3 * D), 30).to(torch.bfloat16)
dctx = _rand((B * N, D), 31, 0.5).to(torch.bfloat16)
ctx_ref, dqkv_ref, lse_ref = _attention_autograd(qkv, dctx, B, H, N, 0.125)
ctx = torch.empty(B * N, D, device="cuda", dtype=torch.bfloat16)
lse = torch.full((B, H, (N + 127) // 128 * 128), float("nan"), device="cuda")
check(lib.vitocm_attention_fwd_lse(eng, ptr(qkv), qkv.stride(0), B, N, ptr(ctx), ctx.stride(0), ptr(lse), cur_stream()))
torch.cuda.synchronize()
dqkv = torch.full((B * N, 3 * D), float("nan"), device="cuda", dtype=torch.bfloat16)
delta = torch.empty(B, H, (N + 127) // 128 * 128, device="cuda")
dqacc = torch.zeros(B * N, D, device="cuda")
check(lib.vitocm_attention_bwd(eng, ptr(qkv), qkv.stride(0), ptr(ctx), ptr(dctx), dctx.stride(0), ptr(lse), ptr(delta), ptr(dqacc),
                               ptr(dqkv), dqkv.stride(0), B, N, cur_stream()))
torch.cuda.synchronize()
for name, lo in (("dq", 0), ("dk", D), ("dv", 2 * D)):
    got, ref = dqkv[:, lo:lo + D].float(), dqkv_ref[:, lo:lo + D]
    e = (got - ref).abs()
    print(f"B{B} H{H} N{N} {name}: max err {e.max().item():.4f} of {ref.abs().max().item():.4f}; per 32-row block:",
          [round(e[r:r + 32].max().item(), 3) for r in range(0, min(B * N, 320), 32)], flush=True)
    if name == "dq":
        print("   per 16-col block:", [round(e[:, c:c + 16].max().item(), 3) for c in range(0, D, 16)], flush=True)
if N >= 128:
    got, ref = dqkv[:, :D].float(), dqkv_ref[:, :D]
    torch.set_printoptions(precision=3, linewidth=200)
    for r in (0, 64, 65, 96):
        print("row", r, "got", got[r, :8].tolist(), "\n       ref", ref[r, :8].tolist())
    # does a wrong row match another reference row?
    for r in (64, 96):
        d = (ref[:128, :64] - got[r:r + 1, :64]).abs().max(dim=1).values
        print("row", r, "closest ref row", int(d.argmin()), float(d.min()))
    print("zeros in got rows 64..127:", float((got[64:128] == 0).float().mean()))
