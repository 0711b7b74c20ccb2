"""Accuracy of the one-kernel block tail (the tensor core adds the MLP output onto the residual rows in TMEM) against the separate
kernels it replaces (fp32 reduce-add in L2), both measured against an fp64 statement with the same 16-bit intermediates.
Usage: python tools/tail_accuracy.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import vitocm_b200 as vob  # noqa: E402
from gpu_util import make_engine, ptr, check, cur_stream  # noqa: E402

lib = vob._lib.load_library()
M, D, Hd = 20000, 384, 1536
for precision, dt in ((2, torch.float16), (0, torch.bfloat16)):
    eng = make_engine(embed_dim=D, heads=6, hidden=Hd, precision=precision)
    for rscale in (1.0, 30.0):
        g = torch.Generator(device="cuda").manual_seed(1)
        rnd = lambda *s, sc=1.0: torch.randn(*s, device="cuda", generator=g) * sc
        ctx = rnd(M, D).to(dt)
        Wp, W1, W2 = rnd(D, D, sc=0.05).to(dt), rnd(Hd, D, sc=0.06).to(dt), rnd(D, Hd, sc=0.03).to(dt)
        bp, b1, b2 = rnd(D, sc=0.1), rnd(Hd, sc=0.2), rnd(D, sc=0.1)
        g2, be2, gn, ben = rnd(D, sc=0.1) + 1, rnd(D, sc=0.1), rnd(D, sc=0.1) + 1, rnd(D, sc=0.1)
        resid = rnd(M, D, sc=rscale)
        # fp64 reference with the 16-bit intermediates of the kernels (norm2 rows, hidden activations)
        x1 = resid.double() + ctx.double() @ Wp.double().T + bp.double()
        xn2 = torch.nn.functional.layer_norm(x1, (D,), g2.double(), be2.double(), 1e-6).to(dt).double()
        hid = torch.nn.functional.gelu(xn2 @ W1.double().T + b1.double()).to(dt).double()
        ref = x1 + hid @ W2.double().T + b2.double()
        xa = resid.clone()
        xn = torch.zeros(M, 2 * D, device="cuda", dtype=dt)
        check(lib.vitocm_block_tail(eng, ptr(ctx), ctx.stride(0), ptr(Wp), Wp.stride(0), ptr(bp), ptr(g2), ptr(be2), ptr(W1), W1.stride(0), ptr(W2),
                                    W2.stride(0), M, D, Hd, ptr(b1), ptr(b2), ptr(xa), ptr(gn), ptr(ben), ptr(xn), xn.stride(0), None, 0, None, None, 0, None, cur_stream()))
        xb = resid.clone()
        yn = torch.zeros(M, 2 * D, device="cuda", dtype=dt)
        check(lib.vitocm_gemm_ln(eng, ptr(ctx), ctx.stride(0), ptr(Wp), Wp.stride(0), M, D, D, ptr(bp), ptr(xb), ptr(g2), ptr(be2), ptr(yn),
                                 yn.stride(0), cur_stream()))
        check(lib.vitocm_mlp_fused(eng, ptr(yn), yn.stride(0), ptr(W1), W1.stride(0), ptr(W2), W2.stride(0), M, D, Hd, ptr(b1), ptr(b2), ptr(xb),
                                   cur_stream()))
        torch.cuda.synchronize()
        for name, x in (("block_tail", xa), ("separate", xb)):
            e = x.double() - ref
            # rows whose norm2 value landed on a different 16-bit neighbour dominate the maximum: report mean / median / signed mean
            print(f"precision={precision} resid_scale={rscale:5.1f} {name:10s}: mean|err|={e.abs().mean().item():.3e} median|err|={e.abs().median().item():.3e} "
                  f"max|err|={e.abs().max().item():.3e} mean(err*sign(x))={(e * torch.sign(ref)).mean().item():+.3e}  (|x| mean {ref.abs().mean().item():.2f})")
