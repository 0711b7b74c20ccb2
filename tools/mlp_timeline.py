"""Per-chunk timeline of the fused MLP kernel (SM clocks) for the leader CTA of pair 0 on its second work item, at full-chip load.
Env: ROWS, PRECISION, VITOCM_FUSE_MLP (1 / 8 epilogue-warp variant), VITOCM_MLP_TL_ITEM."""
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
stamps = torch.zeros(2, 16, 8, dtype=torch.int64, device="cuda")
for _ in range(3):
    check(lib.vitocm_mlp_fused_timeline(eng, ptr(A), A.stride(0), ptr(W1), W1.stride(0), ptr(W2), W2.stride(0), M, D, Hd, ptr(b1), ptr(b2),
                                        ptr(X), ptr(stamps), cur_stream()))
torch.cuda.synchronize()
s = stamps.cpu()
t0 = int(s[s > 0].min())
rel = lambda v: int(v) - t0 if v > 0 else -1
print("epilogue warp 0:  wait fc1 | fc1 done | in regs | gelu done | H free | handed      ||  MMA thread: acc free | fc1 issued | wait gelu | gelu ready | fc2 issued")
for c in range(Hd // 128):
    e = [rel(s[0, c, k]) for k in range(6)]
    m = [rel(s[1, c, k]) for k in range(5)]
    print(f" c={c:2d}  " + " ".join(f"{v:7d}" for v in e) + "   ||  " + " ".join(f"{v:7d}" for v in m))
print("item epilogue: wait OUT", rel(s[0, 15, 0]), "OUT complete", rel(s[0, 15, 1]), "done", rel(s[0, 15, 2]))
