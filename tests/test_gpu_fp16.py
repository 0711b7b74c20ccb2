"""GPU (B200): the fp16 operand format (VITOCM_FP16 engines) and the act-split MLP schedule (vitocm_set_layer_mode 1), kernel by
kernel against plain fp32 torch statements of the same ops and, at model level, against the reference's goldens / the CPU oracle.

fp16 engines keep every 16-bit tensor-core operand as IEEE half (11 significand bits) instead of bf16 (8): same MMA rate, 8x less
rounding error; conversions saturate at 65504.  Tolerances: fp16 outputs 2^-11 relative (1e-3 with head room); CLS rows 1e-3
relative (the fp32 bar of the north star; measured 2e-4)."""
import numpy as np
import pytest
import torch

import vitocm_b200 as vob
from conftest import load_golden
from gpu_util import attention, attention_reference, build_model, gemm, make_engine
from oracle import post_oracle as PO
from oracle import vit_oracle as VO
from vitocm_b200._lib import check, cur_stream, ptr

pytestmark = pytest.mark.gpu
FP16 = 2


@pytest.fixture(autouse=True)
def _no_tf32():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield


@pytest.fixture(scope="module")
def engine():
    return make_engine(precision=FP16)


def _rand(shape, seed, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randn(shape, generator=g, device="cuda") * scale


def split_f16(x: torch.Tensor) -> torch.Tensor:
    hi = x.half()
    lo = (x - hi.float()).half()
    return torch.cat([hi, lo], dim=1).contiguous()


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (785, 1152, 384), (785, 384, 1536), (200, 192, 192), (70000, 1152, 384)])
def test_gemm_fp16_operands_all_epilogues(engine, M, N, K):
    A = _rand((M, K), 1).half()
    B = _rand((N, K), 2, 0.05).half()
    bias = _rand((N,), 3, 0.1)
    prod = A.float() @ B.float().T + bias
    out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.float16)
    gemm(engine, A, B, M, N, K, 0, 0, bias, out, N)
    assert (out.float() - prod).abs().max().item() <= 1e-3 * prod.abs().max().item() + 1e-4
    gemm(engine, A, B, M, N, K, 0, 1, bias, out, N)              # five-coefficient sigmoid-form GELU
    ref = torch.nn.functional.gelu(prod)
    assert (out.float() - ref).abs().max().item() <= 1e-3 * ref.abs().max().item() + 2e-5
    y = torch.full((M, N), float("nan"), device="cuda")
    gemm(engine, A, B, M, N, K, 0, 3, bias, y, N)
    assert (y - prod).abs().max().item() <= 2e-5 * prod.abs().max().item() + 1e-6
    resid = _rand((M, N), 4)
    x = resid.clone()
    gemm(engine, A, B, M, N, K, 0, 2, bias, x, N)
    assert (x - (resid + prod)).abs().max().item() <= 2e-5 * (resid + prod).abs().max().item() + 1e-6


def test_gelu5_epilogue_error_is_below_fp16_rounding(engine):
    """gelu(x) over a grid of pre-activations through the fc1 epilogue: A = x (one-hot selection), bias 0."""
    M, K, N = 4096, 64, 64
    xs = torch.linspace(-9, 9, M * N, device="cuda").reshape(M, N)
    # out[m, n] = sum_k A[m, k] * B[n, k]: A = per-row scale, B = identity block -> the epilogue sees A[m, n] itself
    A = xs[:, :K].contiguous()
    hi = A.half()
    B = torch.eye(N, K, device="cuda").half()
    out = torch.empty(M, N, device="cuda", dtype=torch.float16)
    gemm(engine, hi, B, M, N, K, 0, 1, None, out, N)
    ref = torch.nn.functional.gelu(hi.double()).float()
    err = (out.float() - ref).abs()
    assert (err <= 6e-6 + 2.0 ** -11 * ref.abs()).all(), err.max().item()


@pytest.mark.parametrize("M,N,K,epi", [(785, 1536, 384, 1), (1000, 384, 1536, 2), (70000, 1536, 384, 1), (66000, 384, 1536, 2), (300, 128, 128, 3)])
def test_gemm_act_split_two_terms(engine, M, N, K, epi):
    """split_in = 2: A is a (hi | lo) pair, B single -> hi*B + lo*B: the activation enters at ~2^-22, the weight at fp16."""
    A32 = _rand((M, K), 11)
    B = _rand((N, K), 12, 0.05).half()
    bias = _rand((N,), 13, 0.1)
    A = split_f16(A32)
    prod = (A32.double() @ B.double().T + bias.double()).float()
    if epi == 1:     # fc1 of an act-split block: gelu, written as a (hi | lo) pair
        out = torch.full((M, 2 * N), float("nan"), device="cuda", dtype=torch.float16)
        gemm(engine, A, B, M, N, K, 2, 1, bias, out, 2 * N, 1, N)
        ref = torch.nn.functional.gelu(prod.double()).float()
        rec = out[:, :N].float() + out[:, N:].float()
        assert (rec - ref).abs().max().item() <= 1e-5 + 2e-5 * ref.abs().max().item()
    elif epi == 2:
        resid = _rand((M, N), 14)
        x = resid.clone()
        gemm(engine, A, B, M, N, K, 2, 2, bias, x, N)
        assert (x - (resid + prod)).abs().max().item() <= 2e-5 * (resid + prod).abs().max().item()
    else:
        y = torch.empty(M, N, device="cuda")
        gemm(engine, A, B, M, N, K, 2, 3, bias, y, N)
        assert (y - prod).abs().max().item() <= 1e-5 * prod.abs().max().item()


def test_gemm_three_term_split_fp16(engine):
    M, N, K = 785, 384, 384
    A32, B32 = _rand((M, K), 21), _rand((N, K), 22, 0.05)
    y = torch.empty(M, N, device="cuda")
    gemm(engine, split_f16(A32), split_f16(B32), M, N, K, 1, 3, None, y, N)
    ref = (A32.double() @ B32.double().T).float()
    assert ((y - ref).abs().max() / ref.abs().max()).item() < 6e-6     # fp32 accumulation over K = 384 is the floor here


@pytest.mark.parametrize("M,N,K", [(1000, 384, 384), (25120, 384, 384), (777, 768, 384)])
def test_gemm_fused_layernorm_fp16(engine, M, N, K):
    lib = vob._lib.load_library()
    A = _rand((M, K), 40).half()
    B = _rand((N, K), 41, 0.05).half()
    bias = _rand((N,), 42, 0.1)
    resid = _rand((M, N), 43) + 0.3
    gamma, beta = _rand((N,), 44) * 0.1 + 1, _rand((N,), 45) * 0.1
    x = resid.clone()
    xn = torch.full((M, 2 * N), float("nan"), device="cuda", dtype=torch.float16)
    check(lib.vitocm_gemm_ln(engine, ptr(A), A.stride(0), ptr(B), B.stride(0), M, N, K, ptr(bias), ptr(x), ptr(gamma), ptr(beta),
                             ptr(xn), xn.stride(0), cur_stream()))
    torch.cuda.synchronize()
    ref_x = resid + A.float() @ B.float().T + bias
    assert (x - ref_x).abs().max().item() <= 2e-5 * ref_x.abs().max().item() + 1e-6
    ref_n = torch.nn.functional.layer_norm(x, (N,), gamma, beta, 1e-6)
    assert (xn[:, :N].float() - ref_n).abs().max().item() <= 3e-3


def test_layernorm_fp16(engine):
    lib = vob._lib.load_library()
    M, D = 1000, 128
    x = _rand((M, D), 14, 3.0) + 0.5
    g, b = _rand((D,), 15) * 0.1 + 1, _rand((D,), 16) * 0.1
    out = torch.empty(M, 2 * D, device="cuda", dtype=torch.float16)
    check(lib.vitocm_layernorm(engine, ptr(x), ptr(g), ptr(b), ptr(out), 2 * D, 1, D, M, cur_stream()))
    torch.cuda.synchronize()
    ref = torch.nn.functional.layer_norm(x, (D,), g, b, 1e-6)
    assert (out[:, :D].float() - ref).abs().max().item() < 3e-3
    assert ((out[:, :D].float() + out[:, D:].float()) - ref).abs().max().item() < 2e-6


@pytest.mark.parametrize("B,H,N", [(1, 2, 17), (2, 2, 65), (1, 2, 129), (2, 2, 300), (4, 6, 785), (7, 2, 200)])
def test_attention_fp16(B, H, N):
    eng = make_engine(embed_dim=64 * H, heads=H, precision=FP16)
    D = 64 * H
    q, k, v = (_rand((B, H, N, 64), s) for s in (20, 21, 22))
    qkv = torch.stack([q, k, v], 0).permute(1, 3, 0, 2, 4).reshape(B * N, 3 * D).half().contiguous()
    s5 = qkv.float().reshape(B, N, 3, H, 64).permute(2, 0, 3, 1, 4)
    ref = attention_reference(s5[0], s5[1], s5[2], 0.125).reshape(B * N, D)
    ctx = torch.full((B * N, D), float("nan"), device="cuda", dtype=torch.float16)
    attention(eng, qkv, B, N, ctx)
    err = (ctx.float() - ref).abs().max().item()
    assert err <= 2e-3 * max(1.0, ref.abs().max().item()), (err, ref.abs().max().item())
    vob._lib.load_library().vitocm_destroy(eng)


@pytest.mark.parametrize("gain", [6.0, 20.0, 60.0])
def test_attention_fp16_rising_logits_stay_finite(gain):
    """fp16 P must stay below 65504: a row that jumps more than 2^14 above its running maximum inside one KV block is redone
    against the raised maximum (bf16 engines only do that beyond 2^60)."""
    B, H, N = 1, 2, 600
    eng = make_engine(embed_dim=64 * H, heads=H, precision=FP16)
    D = 64 * H
    q, k, v = (_rand((B, H, N, 64), s) for s in (60, 61, 62))
    ramp = torch.linspace(0.2, 1.0, N, device="cuda").view(1, 1, N, 1)
    k = k * ramp * gain
    qkv = torch.stack([q, k, v], 0).permute(1, 3, 0, 2, 4).reshape(B * N, 3 * D)
    qkv = qkv.clamp(-60000, 60000).half().contiguous()
    s5 = qkv.float().reshape(B, N, 3, H, 64).permute(2, 0, 3, 1, 4)
    ref = attention_reference(s5[0], s5[1], s5[2], 0.125).reshape(B * N, D)
    ctx = torch.full((B * N, D), float("nan"), device="cuda", dtype=torch.float16)
    attention(eng, qkv, B, N, ctx)
    assert torch.isfinite(ctx.float()).all()
    assert (ctx.float() - ref).abs().max().item() <= 4e-3 * max(1.0, ref.abs().max().item())
    vob._lib.load_library().vitocm_destroy(eng)


def rel_err(a, b):
    return float((np.abs(a - b) / np.abs(b)).max())


TINY = VO.ViTConfig(embed_dim=128, depth=3, num_heads=2, patch_size=8, img_size=32)


@pytest.mark.parametrize("precision", ["fp16", "fp16+mlp2", "fp16+mlp2:1", "bf16+mlp2"])
def test_tiny_model_matches_reference_goldens_fp16(precision):
    sd = VO.randomize_affine(VO.init_state_dict(TINY, seed=7), seed=8)
    g = load_golden("tiny_vit.npz")
    m = build_model(TINY, sd, precision, chunk_tiles=2)
    tol = 2e-2 if precision.startswith("bf16") else 1e-3
    for name in ("a", "b", "c"):
        x = torch.from_numpy(g[f"{name}/x"]).cuda()
        ref_attn = g[f"{name}/attn"]
        rows = m.cls_attention_rows(x).cpu().numpy()
        assert rel_err(rows, ref_attn[:, :, 0, :]) <= tol, (name, rel_err(rows, ref_attn[:, :, 0, :]))
        full = m.get_last_selfattention(x).cpu().numpy()
        assert rel_err(full, ref_attn) <= tol
        feat, attns, qkvs = m.get_intermediate_feat(x, n=1)
        f = feat[0].materialize().cpu().numpy()
        assert np.abs(f - g[f"{name}/feat"]).max() <= (6e-2 if precision.startswith("bf16") else 6e-3)


def test_vits8_precision_ladder():
    """ViT-S/8, 8 synthetic tiles: CLS-row error against the CPU oracle falls along bf16 > fp16 > fp16+mlp2 > fp32, and fp16 stays
    inside the north star's fp32 bar (1e-3 relative)."""
    cfg = VO.ViTConfig(**VO.VIT_SMALL)
    sd = VO.randomize_affine(VO.init_state_dict(cfg, seed=0), seed=1, scale=0.02)
    x = torch.cat([VO.synthetic_tile(224, seed=100 + i, batch=1) for i in range(4)])
    ref = VO.cls_attention_rows(sd, cfg, x).numpy()
    rms = {}
    for precision in ("bf16", "fp16", "fp16+mlp2", "fp32"):
        rows = build_model(cfg, sd, precision).cls_attention_rows(x.cuda()).cpu().numpy()
        rms[precision] = float(np.sqrt((((rows - ref) / ref) ** 2).mean()))
        if precision != "bf16":
            assert rel_err(rows, ref) <= 1e-3, (precision, rel_err(rows, ref))
    print("\nCLS-row rms relative error:", rms)
    assert rms["bf16"] > 4 * rms["fp16"] > 4 * rms["fp32"]
    assert rms["fp16+mlp2"] < 0.8 * rms["fp16"]


@pytest.mark.parametrize("precision", ["fp16", "bf16"])
def test_block_tail_folded_layernorm_affine_equals_unfolded(precision, monkeypatch):
    """vitocm_finalize_weights folds norm2 into W1 / b1 and the next norm1 into Wqkv / bqkv for the block-tail kernel
    (VITOCM_TAIL_FOLD, read at vitocm_create).  With strongly non-trivial gamma / beta both forms stay inside the precision's
    CLS-row bar against the CPU oracle and agree with each other far inside it."""
    tol = 2e-2 if precision == "bf16" else 1e-3
    for cfg, x in ((TINY, torch.from_numpy(load_golden("tiny_vit.npz")["a/x"])),
                   (VO.ViTConfig(**VO.VIT_SMALL), VO.synthetic_tile(224, seed=321, batch=1))):
        sd = VO.randomize_affine(VO.init_state_dict(cfg, seed=3), seed=4, scale=0.3 if cfg is TINY else 0.02)
        ref = VO.cls_attention_rows(sd, cfg, x).numpy()
        ref_feat = VO.forward_feats(sd, cfg, x).numpy()   # O(1) dynamic range: a dropped beta or gamma shows here at once
        atol = 6e-2 if precision == "bf16" else 6e-3
        rows, feats = {}, {}
        for fold in ("1", "0"):
            monkeypatch.setenv("VITOCM_TAIL_FOLD", fold)
            m = build_model(cfg, sd, precision, chunk_tiles=2)
            rows[fold] = m.cls_attention_rows(x.cuda()).cpu().numpy()             # chained blocks: fold2 and foldn (QKV rides along)
            feats[fold] = m.forward_feats(x.cuda()).cpu().numpy()                 # block by block: fold2 only
            assert rel_err(rows[fold], ref) <= tol, (fold, rel_err(rows[fold], ref))
            assert np.abs(feats[fold] - ref_feat).max() <= atol, (fold, np.abs(feats[fold] - ref_feat).max())
        assert not np.array_equal(rows["1"], rows["0"])   # the knob really switches the path
        assert rel_err(rows["1"], rows["0"]) <= tol / 2, rel_err(rows["1"], rows["0"])
        print(f"\n{precision} D={cfg.embed_dim}: rows vs oracle folded {rel_err(rows['1'], ref):.2e} / plain {rel_err(rows['0'], ref):.2e}; "
              f"feat folded {np.abs(feats['1'] - ref_feat).max():.2e} / plain {np.abs(feats['0'] - ref_feat).max():.2e}")
