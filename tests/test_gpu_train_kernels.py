"""GPU (B200): the training-step kernels (weight-gradient GEMM, flash-attention backward, fused clip + AdamW)
against plain fp32 torch statements of the same ops, called through the C ABI."""
import pytest
import torch

import vitocm_b200 as vob
from gpu_util import make_engine
from vitocm_b200._lib import check, cur_stream, ptr

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _no_tf32():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield


@pytest.fixture(scope="module")
def engine():
    return make_engine()


def _rand(shape, seed, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randn(shape, generator=g, device="cuda") * scale


# (M tokens, R = rows of dW, C = columns of dW): every tile configuration (C % 384, 256, 192, 128, 64), ragged M,
# R not a multiple of 128, token splits > 1 and == 1
@pytest.mark.parametrize("M,R,C", [(64, 128, 64), (100, 128, 128), (785, 384, 384), (1570, 1152, 384), (3000, 384, 1536),
                                   (785, 192, 384), (1568, 384, 192), (333, 64, 256), (20000, 128, 128), (50, 72, 64)])
def test_wgrad(engine, M, R, C):
    lib = vob._lib.load_library()
    G = _rand((M, R), 1, 0.5).to(torch.bfloat16)
    A = _rand((M, C), 2).to(torch.bfloat16)
    dW0 = _rand((R, C), 3)
    dW = dW0.clone()
    db0 = _rand((R,), 6)
    db = db0.clone()
    check(lib.vitocm_wgrad(engine, ptr(G), G.stride(0), ptr(A), A.stride(0), M, R, C, ptr(dW), ptr(db), cur_stream()))
    torch.cuda.synchronize()
    ref = dW0.double() + G.double().T @ A.double()
    err = (dW.double() - ref).abs().max().item()
    assert err <= 2e-5 * ref.abs().max().item() + 1e-4, (err, ref.abs().max().item())
    # the bias gradient (column sums of G) from the all-ones MMA of the same pass
    ref_b = db0.double() + G.double().sum(0)
    assert (db.double() - ref_b).abs().max().item() <= 2e-5 * ref_b.abs().max().item() + 1e-4


def test_wgrad_strided_operands(engine):
    """operands are column slices of wider activations (the qkv gradient is read out of [M][3D])"""
    lib = vob._lib.load_library()
    M, R, C = 785, 256, 128
    Gw = _rand((M, 3 * R), 4, 0.5).to(torch.bfloat16)
    Aw = _rand((M, 2 * C), 5).to(torch.bfloat16)
    G, A = Gw[:, R:2 * R], Aw[:, C:]
    dW = torch.zeros(R, C, device="cuda")
    check(lib.vitocm_wgrad(engine, G.data_ptr(), Gw.stride(0), A.data_ptr(), Aw.stride(0), M, R, C, ptr(dW), None, cur_stream()))
    torch.cuda.synchronize()
    ref = G.double().T @ A.double()
    assert (dW.double() - ref).abs().max().item() <= 2e-5 * ref.abs().max().item() + 1e-4


def _attention_autograd(qkv_bf16, dctx_bf16, B, H, N, scale):
    """fp32 torch autograd of vit.py:83-87 on the bf16-rounded inputs -> (ctx, dqkv) in the [B*N][3D] layout"""
    D = 64 * H
    src = qkv_bf16.float().reshape(B, N, 3, H, 64).permute(2, 0, 3, 1, 4).contiguous().requires_grad_(True)
    q, k, v = src[0], src[1], src[2]
    a = ((q @ k.transpose(-2, -1)) * scale).softmax(dim=-1)
    ctx = (a @ v).transpose(1, 2).reshape(B * N, D)
    ctx.backward(dctx_bf16.float())
    dqkv = src.grad.permute(1, 3, 0, 2, 4).reshape(B * N, 3 * D)
    lse2 = torch.logsumexp((q @ k.transpose(-2, -1)) * scale, dim=-1).detach() * 1.4426950408889634   # [B,H,N]
    return ctx.detach(), dqkv, lse2


@pytest.mark.parametrize("B,H,N", [(1, 2, 17), (1, 2, 128), (2, 2, 129), (2, 2, 300), (1, 6, 785), (3, 2, 37)])
def test_attention_backward(B, H, N):
    lib = vob._lib.load_library()
    eng = make_engine(embed_dim=64 * H, heads=H, precision=0)
    D = 64 * H
    qkv = _rand((B * N, 3 * D), 30).to(torch.bfloat16)
    dctx = _rand((B * N, D), 31, 0.5).to(torch.bfloat16)
    ctx_ref, dqkv_ref, lse_ref = _attention_autograd(qkv, dctx, B, H, N, 0.125)
    ctx = torch.empty(B * N, D, device="cuda", dtype=torch.bfloat16)
    npad = (N + 127) // 128 * 128
    lse = torch.full((B, H, npad), float("nan"), device="cuda")
    check(lib.vitocm_attention_fwd_lse(eng, ptr(qkv), qkv.stride(0), B, N, ptr(ctx), ctx.stride(0), ptr(lse), cur_stream()))
    torch.cuda.synchronize()
    assert (lse[:, :, :N] - lse_ref).abs().max().item() < 2e-2
    assert torch.isinf(lse[:, :, N:]).all()          # pad rows of the 128-query tiles carry +inf
    assert (ctx.float() - ctx_ref).abs().max().item() < 2e-2 * max(1.0, ctx_ref.abs().max().item())
    dqkv = torch.full((B * N, 3 * D), float("nan"), device="cuda", dtype=torch.bfloat16)
    delta = torch.empty(B, H, npad, device="cuda")
    dqacc = torch.zeros(B * N, D, device="cuda")
    check(lib.vitocm_attention_bwd(eng, ptr(qkv), qkv.stride(0), ptr(ctx), ptr(dctx), dctx.stride(0), ptr(lse), ptr(delta), ptr(dqacc),
                                   ptr(dqkv), dqkv.stride(0), B, N, cur_stream()))
    torch.cuda.synchronize()
    assert torch.isfinite(dqkv.float()).all()
    assert (dqacc == 0).all(), "dQ accumulator must be left zeroed"
    for name, lo in (("dq", 0), ("dk", D), ("dv", 2 * D)):
        got, ref = dqkv[:, lo:lo + D].float(), dqkv_ref[:, lo:lo + D]
        err = (got - ref).abs().max().item()
        assert err <= 3e-2 * ref.abs().max().item() + 1e-3, (name, err, ref.abs().max().item())
    lib.vitocm_destroy(eng)


def test_clip_and_adamw_match_torch():
    lib = vob._lib.load_library()
    n = 100003
    p0 = _rand((n,), 40)
    decay = (torch.arange(n, device="cuda") % 3 != 0).to(torch.uint8)
    ref_p = torch.nn.Parameter(p0.clone())
    groups = [{"params": [ref_p]}]
    # torch reference: one parameter per decay class would need a split; emulate with two tensors
    pa, pb = torch.nn.Parameter(p0[decay.bool()].clone()), torch.nn.Parameter(p0[~decay.bool()].clone())
    opt = torch.optim.AdamW([{"params": [pa]}, {"params": [pb], "weight_decay": 0.0}], lr=5e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.05)
    p, m, v = p0.clone(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    ss = torch.zeros(1, device="cuda", dtype=torch.float64)
    for step in range(1, 4):
        g = _rand((n,), 50 + step, 0.1 * step)
        pa.grad, pb.grad = g[decay.bool()].clone(), g[~decay.bool()].clone()
        total = torch.nn.utils.clip_grad_norm_([pa, pb], 5.0)
        opt.step()
        gg = g.clone()
        check(lib.vitocm_grad_sumsq(ptr(gg), n, ptr(ss), cur_stream()))
        check(lib.vitocm_adamw_step(ptr(p), ptr(gg), ptr(m), ptr(v), ptr(decay), n, 5e-4, 0.9, 0.999, 1e-8, 0.05, step, 5.0, 1.0, ptr(ss),
                                    cur_stream()))
        torch.cuda.synchronize()
        assert abs(ss.sqrt().item() - total.item()) <= 1e-5 * total.item()
        assert (p[decay.bool()] - pa.detach()).abs().max().item() < 2e-6
        assert (p[~decay.bool()] - pb.detach()).abs().max().item() < 2e-6
    del groups, ref_p
