"""Per-chunk timeline of the fused MLP kernel (SM clocks) for the leader CTA of cluster 0 on its second work item, at full-chip load.
Env: ROWS, PRECISION, VITOCM_FUSE_MLP (cluster size 4 / 2), VITOCM_MLP_TL_ITEM, VITOCM_MLP_DEBUG."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import vitocm_b200 as vob
from vitocm_b200._lib import check, cur_stream, ptr
from gpu_util import make_engine

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
stamps = torch.zeros(64, dtype=torch.int64, device="cuda")
for _ in range(3):
    check(lib.vitocm_mlp_fused_timeline(eng, ptr(A), A.stride(0), ptr(W1), W1.stride(0), ptr(W2), W2.stride(0), M, D, Hd, ptr(b1), ptr(b2),
                                        ptr(X), ptr(stamps), cur_stream()))
torch.cuda.synchronize()
s = [int(v) for v in stamps.cpu()]
t0 = s[60]
rel = lambda v: (v - t0) & 0xffffffff
print(f"fused MLP timeline (cluster {os.environ.get('VITOCM_FUSE_MLP', '4')}, debug {os.environ.get('VITOCM_MLP_DEBUG', '0')}); clocks since the item's A tile landed")
print("        epilogue warp 0: fc1(c) complete | gelu arithmetic done | gelu(c) handed over   ||   MMA thread: fc1(c) issued | gelu(c) available")
for c in range(Hd // 128):
    print(f" c={c:2d}   {rel(s[3*c]):8d} {rel(s[3*c+1]):8d} {rel(s[3*c+2]):8d}   ||   {rel(s[36+2*c]):8d} {rel(s[36+2*c+1]):8d}")
print(f"OUT complete {rel(s[61])}, item epilogue done {rel(s[62])}")
