"""Fused MLP kernel against the fc1 + fc2 GEMM pair it replaces, at the bench chunk (M = 175 x 785 rows of ViT-S): device time per
launch.  Env: ROWS, PRECISION (0 bf16 / 2 fp16), VITOCM_FUSE_MLP (cluster size 4 or 2), VITOCM_MLP_STAGGER."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import vitocm_b200 as vob
from gpu_util import make_engine
from vitocm_b200._lib import check, cur_stream, ptr
M = int(os.environ.get("ROWS", str(175 * 785)))
PREC = int(os.environ.get("PRECISION", "0"))
D, Hd = 384, 1536
dt = torch.float16 if PREC == 2 else torch.bfloat16
eng = make_engine(embed_dim=D, heads=6, hidden=Hd, precision=PREC)
lib = vob._lib.load_library()
A = (torch.randn(M, 2 * D, device="cuda") * 0.5).to(dt)
W1 = (torch.randn(Hd, D, device="cuda") * 0.05).to(dt)
W2 = (torch.randn(D, Hd, device="cuda") * 0.03).to(dt)
b1, b2 = torch.randn(Hd, device="cuda") * 0.1, torch.randn(D, device="cuda") * 0.1
X = torch.zeros(M, D, device="cuda")
HID = torch.empty(M, Hd, device="cuda", dtype=dt)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def fused():
    check(lib.vitocm_mlp_fused(eng, ptr(A), A.stride(0), ptr(W1), W1.stride(0), ptr(W2), W2.stride(0), M, D, Hd, ptr(b1), ptr(b2), ptr(X), cur_stream()))
def separate():
    check(lib.vitocm_gemm(eng, ptr(A), A.stride(0), ptr(W1), W1.stride(0), M, Hd, D, 0, 1, ptr(b1), ptr(HID), Hd, 0, 0, cur_stream()))
    check(lib.vitocm_gemm(eng, ptr(HID), HID.stride(0), ptr(W2), W2.stride(0), M, D, Hd, 0, 2, ptr(b2), ptr(X), D, 0, 0, cur_stream()))
for name, fn in (("fused", fused), ("fc1+fc2", separate)):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(30): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 30
    print(f"mlp {name} prec={PREC} cluster={os.environ.get('VITOCM_FUSE_MLP','4')} stagger={os.environ.get('VITOCM_MLP_STAGGER','40000')} M={M}: {ms*1e3:.1f} us/launch, {4*M*D*Hd/ms/1e9:.0f} TFLOP/s")
