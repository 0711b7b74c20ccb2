"""The four GEMM shapes of a ViT-S block at the bench chunk (M = 32 x 785): device time per launch.
Env: VITOCM_GEMM_RESIDENT=0/1, VITOCM_GEMM_DEBUG=0/1/2 (1: epilogue drains TMEM only, 2: epilogue signals only)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import vitocm_b200 as vob
from gpu_util import gemm, make_engine
from vitocm_b200._lib import check, cur_stream, ptr
M = int(os.environ.get("ROWS", str(32 * 785)))
eng = make_engine()
lib = vob._lib.load_library()
tag = f"res={os.environ.get('VITOCM_GEMM_RESIDENT','1')} dbg={os.environ.get('VITOCM_GEMM_DEBUG','0')}"
for name, N, K, epi in (("qkv", 1152, 384, 0), ("proj", 384, 384, 2), ("fc1", 1536, 384, 1), ("fc2", 384, 1536, 2)):
    A = (torch.randn(M, K, device="cuda") * 0.5).to(torch.bfloat16)
    B = (torch.randn(N, K, device="cuda") * 0.05).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda") * 0.1
    out = torch.zeros(M, N, device="cuda", dtype=torch.float32 if epi >= 2 else torch.bfloat16)
    def run(n):
        for _ in range(n):
            check(lib.vitocm_gemm(eng, ptr(A), A.stride(0), ptr(B), B.stride(0), M, N, K, 0, epi, ptr(bias), ptr(out), N, 0, 0, cur_stream()))
    run(5); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(40); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 40
    line = f"{tag} {name}: M={M} N={N} K={K}: {ms*1e3:.1f} us, {2*M*N*K/ms/1e9:.0f} TFLOP/s"
    if os.environ.get("CUBLAS_YARDSTICK"):
        Bt = B.t().contiguous()
        for _ in range(5): torch.matmul(A, Bt)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(40): torch.matmul(A, Bt)
        e1.record(); torch.cuda.synchronize()
        ms2 = e0.elapsed_time(e1) / 40
        line += f" | cuBLAS plain bf16 matmul (no epilogue) yardstick: {ms2*1e3:.1f} us, {2*M*N*K/ms2/1e9:.0f} TFLOP/s"
    print(line)
