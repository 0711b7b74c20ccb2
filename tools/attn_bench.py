"""Attention kernel alone at the bench shape (32 tiles x 6 heads x 785 tokens): device time per launch."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import vitocm_b200 as vob
from vitocm_b200._lib import check, cur_stream, ptr
from gpu_util import make_engine
B, H, N = int(os.environ.get("TILES", "32")), int(os.environ.get("HEADS", "6")), int(os.environ.get("TOKENS", "785"))
D = 64 * H
PREC = int(os.environ.get("PRECISION", "0"))      # 0 = bf16, 2 = fp16
DT = torch.float16 if PREC == 2 else torch.bfloat16
eng = make_engine(embed_dim=D, heads=H, precision=PREC)
qkv = torch.randn(B * N, 3 * D, device="cuda").to(DT)
ctx = torch.empty(B * N, D, device="cuda", dtype=DT)
lib = vob._lib.load_library()
def run(n):
    for _ in range(n):
        check(lib.vitocm_attention(eng, ptr(qkv), qkv.stride(0), B, N, ptr(ctx), ctx.stride(0), cur_stream()))
run(5); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(50); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 50
print(f"track_max={os.environ.get('VITOCM_ATTN_TRACK_MAX','0')} precision={PREC} B={B} H={H} N={N}: {ms*1e3:.1f} us/launch, {4*B*H*N*N*64/ms/1e9:.1f} TFLOP/s")
