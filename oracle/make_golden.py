"""Generate tests/golden/*.npz by running the REFERENCE's own Python code (imported read-only
from /root/reference) on seeded synthetic inputs, and check the oracle restatement against it.

Run in the build container only (the GPU box has no /root/reference):
    PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_golden

What comes from where
  * dino/vision_transformer.py, dino/utils.py : imported as-is.
  * model.py : imported with a 3-line stub for `timm.models.layers.trunc_normal_` (timm absent).
  * utils.py / sw_processing.py : cannot be imported (matplotlib, scikit-image absent);
    the needed function bodies are exec'd from their source line ranges into a namespace with
    numpy + cv2 (+ a stub for skimage.filters.threshold_otsu -> cv2 Otsu, off the hot path).
Nothing from the reference is copied into this repository: only its OUTPUTS are stored.
"""
from __future__ import annotations

import os
import sys
import types
from functools import partial

import numpy as np
import torch

REF = "/root/reference/Self-supervised_segmentation"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _import_reference():
    sys.dont_write_bytecode = True
    if REF not in sys.path:
        sys.path.insert(0, REF)
    timm = types.ModuleType("timm")
    timm_models = types.ModuleType("timm.models")
    timm_layers = types.ModuleType("timm.models.layers")
    timm_layers.trunc_normal_ = torch.nn.init.trunc_normal_
    sys.modules.setdefault("timm", timm)
    sys.modules.setdefault("timm.models", timm_models)
    sys.modules.setdefault("timm.models.layers", timm_layers)
    import dino.vision_transformer as vits   # noqa
    import model as ref_model                # noqa
    return vits, ref_model


def _exec_lines(path: str, ranges, ns: dict):
    with open(path) as f:
        lines = f.readlines()
    src = "".join("".join(lines[a - 1:b]) + "\n" for a, b in ranges)
    exec(compile(src, path, "exec"), ns)
    return ns


def _ref_post_namespaces():
    import cv2

    class _Filters:
        @staticmethod
        def threshold_otsu(img):
            return cv2.threshold(np.asarray(img), 0, 255, cv2.THRESH_BINARY + cv2.THRESH_OTSU)[0]

    skimage = types.SimpleNamespace(filters=_Filters)
    base = dict(np=np, cv2=cv2, os=os, plt=None, skimage=skimage, create_dir=lambda p: None, nn=torch.nn, torch=torch)
    ns_utils = _exec_lines(os.path.join(REF, "utils.py"), [(55, 115), (229, 235), (304, 317)], dict(base))
    ns_sw = _exec_lines(os.path.join(REF, "sw_processing.py"), [(29, 81), (113, 163)], dict(base))
    return ns_utils, ns_sw


def checksum(sd):
    return {k: np.array([float(v.double().sum()), float(v.double().abs().sum())]) for k, v in sd.items()}


def main():
    import cv2
    from PIL import Image
    from oracle import vit_oracle as VO
    from oracle import post_oracle as PO

    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    vits, ref_model = _import_reference()
    ns_utils, ns_sw = _ref_post_namespaces()
    report = []

    # ---------------------------------------------------------------- known answers (log)
    m = vits.vit_small(patch_size=8, num_classes=0)
    n_vits = sum(p.numel() for p in m.parameters())
    assert n_vits == 21670272, n_vits                      # SSS/output/log_rank0.txt:5570
    enc = ref_model.VisionTransformerForSimMIM(patch_size=8, embed_dim=384, depth=12, num_heads=6, mlp_ratio=4,
                                               img_size=[224], qkv_bias=True,
                                               norm_layer=partial(torch.nn.LayerNorm, eps=1e-6))
    mim = ref_model.MIM(encoder=enc, encoder_stride=8)
    n_mim = sum(p.numel() for p in mim.parameters())
    assert n_mim == 21744576, n_mim                        # SSS/output/log_rank0.txt:9746
    cfg_s = VO.ViTConfig(**VO.VIT_SMALL)
    assert VO.param_count(VO.init_state_dict(cfg_s)) == n_vits
    report.append(f"param counts ok: {n_vits} / {n_mim}")

    # ---------------------------------------------------------------- tiny model, several shapes
    tiny = VO.ViTConfig(embed_dim=128, depth=3, num_heads=2, patch_size=8, img_size=32)
    sd_t = VO.randomize_affine(VO.init_state_dict(tiny, seed=7), seed=8)
    ref_t = vits.VisionTransformer(img_size=[32], patch_size=8, embed_dim=128, depth=3, num_heads=2, mlp_ratio=4,
                                   qkv_bias=True, norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), num_classes=0)
    ref_t.load_state_dict(sd_t, strict=True)
    ref_t.eval()
    g = {}
    for k, v in checksum(sd_t).items():
        g["wsum/" + k] = v
    for name, (B, S) in {"a": (2, 32), "b": (1, 48), "c": (3, 64)}.items():
        x = VO.synthetic_tile(S, seed=100 + S, batch=B)
        with torch.no_grad():
            feat, attns, qkvs = ref_t.get_intermediate_feat(x, n=1)
            last = ref_t.get_last_selfattention(x)
            ff = ref_t.forward_feats(x)
            cls = ref_t(x)
        assert torch.equal(attns[0], last)
        # oracle restatement vs reference
        o_last = VO.get_last_selfattention(sd_t, tiny, x)
        o_feat, o_attn, o_qkv = VO.get_intermediate_feat(sd_t, tiny, x)
        for a, b, what in [(o_last, last, "last"), (o_feat[0], feat[0], "feat"), (o_attn[0], attns[0], "attn"),
                           (o_qkv[0], qkvs[0], "qkv"), (VO.forward_feats(sd_t, tiny, x), ff, "ff")]:
            err = (a - b).abs().max().item()
            assert err < 1e-5, (name, what, err)
        g[f"{name}/x"] = x.numpy()
        g[f"{name}/attn"] = last.numpy()
        g[f"{name}/feat"] = feat[0].numpy()
        g[f"{name}/qkv"] = qkvs[0].numpy()
        g[f"{name}/cls"] = cls.numpy()
    np.savez_compressed(os.path.join(OUT, "tiny_vit.npz"), **g)
    report.append("tiny_vit.npz: oracle == reference (<1e-5) on 3 shapes")

    # ---------------------------------------------------------------- ViT-S/8, one 224 tile (config 1)
    sd_s = VO.randomize_affine(VO.init_state_dict(cfg_s, seed=0), seed=1, scale=0.02)
    ref_s = vits.vit_small(patch_size=8, num_classes=0)
    ref_s.load_state_dict(sd_s, strict=True)
    ref_s.eval()
    x = VO.synthetic_tile(224, seed=1234, batch=1)
    with torch.no_grad():
        feat, attns, qkvs = ref_s.get_intermediate_feat(x, n=1)
    rows_ref = attns[0][:, :, 0, :].contiguous()
    rows_orc = VO.cls_attention_rows(sd_s, cfg_s, x)
    err = ((rows_ref - rows_orc).abs() / rows_ref.abs()).max().item()
    assert err < 1e-4, err
    # reference post chain (eval.py:136-173), exec'd reference functions
    att_resp, nh = ns_utils["compute_attention"](attns, 0, 28, 28, 8)
    avg = np.mean(att_resp, axis=0)
    avg = cv2.resize(avg, (avg.shape[1] // 8, avg.shape[0] // 8))
    avg = cv2.resize(avg, (224, 224), interpolation=cv2.INTER_LINEAR)
    import torchvision.transforms as T
    pil = T.ToPILImage()(x.squeeze(0)).convert("L")
    th, th2, th3 = ns_utils["threshold"](pil, avg, save=False)
    o_th, o_th2, o_th3, o_res, o_att = PO.eval_tile(rows_ref[0].numpy(), x[0, 0].numpy(), 8)
    for a, b, what in [(o_th, th, "th"), (o_th2, th2, "th2"), (o_th3, th3, "th3")]:
        agree = float((a == b).mean())
        assert agree >= 0.9999, (what, agree)
    att_o = PO.tile_attention_map(rows_ref[0].numpy(), 224, 8)
    assert np.abs(att_o - avg).max() <= 1e-6 * np.abs(avg).max(), np.abs(att_o - avg).max()
    g = {"x_seed": np.array(1234), "cls_rows": rows_ref.numpy(), "att_map": avg.astype(np.float32),
         "th": th, "th2": th2, "th3": th3, "feat_cls": feat[0][:, 0].numpy()}
    for k, v in checksum(sd_s).items():
        g["wsum/" + k] = v
    np.savez_compressed(os.path.join(OUT, "vits8_tile.npz"), **g)
    report.append(f"vits8_tile.npz: cls rows rel err oracle-vs-ref {err:.2e}; masks agree")

    # ---------------------------------------------------------------- Otsu / resize restatements vs cv2
    rng = np.random.RandomState(0)
    otsu_imgs, otsu_t = [], []
    for k in range(64):
        kind = k % 4
        if kind == 0:
            img = rng.randint(0, 256, (40, 40)).astype(np.uint8)
        elif kind == 1:
            img = np.clip(rng.normal(60, 20, (40, 40)), 0, 255).astype(np.uint8)
        elif kind == 2:
            img = np.where(rng.rand(40, 40) < 0.3, rng.randint(150, 256, (40, 40)), rng.randint(0, 80, (40, 40))).astype(np.uint8)
        else:
            img = np.full((40, 40), rng.randint(0, 256), np.uint8)
            img[: k % 7] = rng.randint(0, 256)
        t, mask = cv2.threshold(img, 0, 255, cv2.THRESH_BINARY + cv2.THRESH_OTSU)
        to, mo = PO.otsu_threshold(img)
        assert int(t) == to and np.array_equal(mask, mo), (k, t, to)
        otsu_imgs.append(img)
        otsu_t.append(int(t))
    small = rng.rand(6, 7, 7).astype(np.float32)
    ups = []
    for s in small:
        u = cv2.resize(s, (56, 56), interpolation=cv2.INTER_LINEAR)
        uo = PO.resize_linear(s, (56, 56))
        assert np.abs(u - uo).max() < 1e-6, np.abs(u - uo).max()
        d = cv2.resize(np.repeat(np.repeat(s, 8, 0), 8, 1), (7, 7))
        assert np.abs(d - s).max() < 1e-6
        ups.append(u)
    np.savez_compressed(os.path.join(OUT, "cv2_ops.npz"), otsu_imgs=np.stack(otsu_imgs), otsu_t=np.array(otsu_t),
                        resize_in=small, resize_out=np.stack(ups))
    report.append("cv2_ops.npz: Otsu restatement == cv2 on 64 images; resize within 1e-6")

    # ---------------------------------------------------------------- stitching (reference loops exec'd)
    g = {}
    cases = {"w32s16n4": (32, 16, 4), "w48s16n3": (48, 16, 3), "w24s8n5": (24, 8, 5), "w32s16n1": (32, 16, 1)}
    for name, (W, S, n) in cases.items():
        tiles = [(rng.rand(W, W) * 255).astype(np.float32) for _ in range(n * n)]
        ref = ns_sw["concat_crops"](tiles, S, W)
        orc = PO.concat_crops_blend(tiles, S, W)
        assert ref.dtype == np.float32 and np.array_equal(ref, orc), name
        P = PO.blend_profiles(n, S, W)
        rec = np.zeros_like(ref, dtype=np.float64)
        for i in range(n):
            for j in range(n):
                rec[i * S:i * S + W, j * S:j * S + W] += P[i, i * S:i * S + W, None] * P[j, None, j * S:j * S + W] * tiles[i * n + j]
        assert np.abs(rec - ref).max() < 1e-3, (name, np.abs(rec - ref).max())
        g[name + "/tiles"] = np.stack(tiles)
        g[name + "/out"] = ref
        E0 = (n - 1) * S + W + 2 * S - 1 if n > 1 else W
        # uint8 image path: sliding_window on a PIL image + uint8 blend (sw_processing.py:223-225)
        img = rng.randint(0, 256, (E0, E0)).astype(np.uint8)
        pil = Image.fromarray(np.stack([img] * 3, -1))
        crops = ns_sw["sliding_window"](pil, S, W)
        ocrops = PO.sliding_window(np.stack([img] * 3, -1), S, W)
        assert len(crops) == len(ocrops) and all(np.array_equal(a, b) for a, b in zip(crops, ocrops)), name
        if len(crops) == n * n:
            st = ns_sw["concat_crops"](crops, S, W)
            so = PO.concat_crops_blend(ocrops, S, W)
            assert np.array_equal(st, so), name
            gray = np.array(Image.fromarray(st).convert("L"))
            assert np.array_equal(gray, st[..., 0])
            g[name + "/img"] = img
            g[name + "/gray_stitched"] = gray
    np.savez_compressed(os.path.join(OUT, "stitch.npz"), **g)
    report.append("stitch.npz: blend restatement bit-equal to reference loops (float32 + uint8), separable profiles ok")

    # ---------------------------------------------------------------- sw threshold (reference exec'd)
    E = 80
    att = (rng.rand(E, E) * 255).astype(np.float32)
    img = np.clip(rng.normal(50, 30, (E, E)), 0, 255).astype(np.uint8)
    th, th2, th3 = ns_sw["threshold"](Image.fromarray(img), att, save=False)
    o = PO.threshold_sw(img, att)
    assert np.array_equal(th, o[0]) and np.array_equal(th3, o[2]) and np.array_equal(th2, o[1])
    th_u, th2_u, th3_u = ns_utils["threshold"](Image.fromarray(img), att, save=False)
    ou = PO.threshold_utils(img, att)
    assert np.array_equal(th_u, ou[0]) and np.array_equal(th3_u, ou[2]) and np.array_equal(th2_u, ou[1])
    np.savez_compressed(os.path.join(OUT, "threshold.npz"), att=att, img=img, sw_th=th, sw_th2=th2, sw_th3=th3,
                        ut_th=th_u, ut_th2=th2_u, ut_th3=th3_u)
    report.append("threshold.npz: both threshold() flavours bit-equal to reference")

    # ---------------------------------------------------------------- MaskGenerator + MIM forward (tiny)
    sys.path.insert(0, REF)
    np.random.seed(0)
    src = _exec_lines(os.path.join(REF, "data.py"), [(163, 186)], dict(np=np, print=lambda *a, **k: None))
    mg = src["MaskGenerator"](input_size=224, mask_patch_size=16, model_patch_size=8, mask_ratio=0.5)
    m_ref = mg()
    m_orc = VO.mask_generator(np.random.RandomState(0), 224, 16, 8, 0.5)
    assert np.array_equal(m_ref, m_orc) and m_ref.shape == (28, 28) and m_ref.sum() == 392
    tiny_m = VO.ViTConfig(embed_dim=128, depth=2, num_heads=2, patch_size=8, img_size=32)
    # the reference wrapper never forwards img_size to the base ctor (model.py:11-13), so pos_embed
    # always has 28*28+1 rows and is bicubically resized when img_size[0] != 224 (model.py:38-39)
    tiny_m_init = VO.ViTConfig(embed_dim=128, depth=2, num_heads=2, patch_size=8, img_size=224)
    sd_m = VO.randomize_affine(VO.init_state_dict(tiny_m_init, seed=11, mim=True), seed=12)
    enc = ref_model.VisionTransformerForSimMIM(patch_size=8, embed_dim=128, depth=2, num_heads=2, mlp_ratio=4,
                                               img_size=[32], qkv_bias=True,
                                               norm_layer=partial(torch.nn.LayerNorm, eps=1e-6))
    mim = ref_model.MIM(encoder=enc, encoder_stride=8)
    enc.load_state_dict(sd_m, strict=True)
    gd = torch.Generator().manual_seed(5)
    dec_w = torch.randn(192, 128, 1, 1, generator=gd) * 0.05
    dec_b = torch.randn(192, generator=gd) * 0.05
    mim.decoder[0].weight.data.copy_(dec_w)
    mim.decoder[0].bias.data.copy_(dec_b)
    xm = VO.synthetic_tile(32, seed=77, batch=4)
    rs = np.random.RandomState(3)
    masks = torch.from_numpy(np.stack([VO.mask_generator(rs, 32, 16, 8, 0.5) for _ in range(4)]))
    loss, x_rec, mk = mim(xm, masks)
    o_loss, o_rec, _ = VO.mim_forward(sd_m, tiny_m, dec_w, dec_b, xm, masks)
    assert abs(loss.item() - o_loss.item()) < 1e-6 and (x_rec - o_rec).abs().max().item() < 1e-5
    g = {"x": xm.numpy(), "mask": masks.numpy(), "loss": np.array(loss.item()), "x_rec": x_rec.detach().numpy(),
         "dec_w": dec_w.numpy(), "dec_b": dec_b.numpy(), "mask224_seed0": m_ref}
    for k, v in checksum(sd_m).items():
        g["wsum/" + k] = v
    np.savez_compressed(os.path.join(OUT, "mim_tiny.npz"), **g)
    report.append("mim_tiny.npz: MaskGenerator + MIM forward oracle == reference")

    report += make_train_golden(ref_model)

    with open(os.path.join(OUT, "README.md"), "w") as f:
        f.write("# Golden fixtures\n\nGenerated by `python -m oracle.make_golden` in the build container from the\n"
                "reference's own code at /root/reference (outputs only; no reference source is stored).\n\n")
        for r in report:
            f.write(f"* {r}\n")
    print("\n".join(report))


def sample(t: torch.Tensor) -> np.ndarray:
    """fixture-size control: 1-D tensors whole, larger ones every 7th element (+ their sums, stored separately)"""
    f = t.detach().reshape(-1)
    return (f if t.dim() <= 1 or f.numel() <= 4096 else f[::7]).numpy().copy()


def make_train_golden(ref_model):
    """Two iterations of the reference training step (SSS/mim.py:172-179) on the tiny MIM: the reference's MIM module,
    its optimizer grouping (SSS/optimizer.py:14-33, exec'd: the module imports nothing exotic), torch's
    clip_grad_norm_ and AdamW -- checked against oracle/train_oracle.py and stored."""
    from oracle import train_oracle as TO
    from oracle import vit_oracle as VO
    tiny_m = VO.ViTConfig(embed_dim=128, depth=2, num_heads=2, patch_size=8, img_size=32)
    tiny_m_init = VO.ViTConfig(embed_dim=128, depth=2, num_heads=2, patch_size=8, img_size=224)
    sd_m = VO.randomize_affine(VO.init_state_dict(tiny_m_init, seed=11, mim=True), seed=12)
    enc = ref_model.VisionTransformerForSimMIM(patch_size=8, embed_dim=128, depth=2, num_heads=2, mlp_ratio=4, img_size=[32],
                                               qkv_bias=True, norm_layer=partial(torch.nn.LayerNorm, eps=1e-6))
    mim = ref_model.MIM(encoder=enc, encoder_stride=8)
    enc.load_state_dict(sd_m, strict=True)
    gd = torch.Generator().manual_seed(5)
    dec_w = torch.randn(192, 128, 1, 1, generator=gd) * 0.05
    dec_b = torch.randn(192, generator=gd) * 0.05
    mim.decoder[0].weight.data.copy_(dec_w)
    mim.decoder[0].bias.data.copy_(dec_b)
    ns = _exec_lines(os.path.join(REF, "optimizer.py"), [(1, 33)], {})
    class _Log:
        def info(self, *a, **k):
            pass
    groups = ns["get_pretrain_param_groups"](mim, _Log(), mim.no_weight_decay(), mim.no_weight_decay_keywords())
    opt = torch.optim.AdamW(groups, eps=1e-8, betas=(0.9, 0.999), lr=5e-4, weight_decay=0.05)   # SSS/optimizer.py:73-75, config.py:98-130
    params = dict(sd_m)
    params["decoder.0.weight"], params["decoder.0.bias"] = dec_w, dec_b
    state = TO.TrainState(params)
    decay_names = {n for n, p in mim.named_parameters() if any(p is q for q in groups[0]["params"])}
    for k, v in params.items():
        assert TO.has_weight_decay(k, v.shape) == (TO.full_name(k) in decay_names), k
    g = {}
    rs = np.random.RandomState(9)
    mim.train()
    for it in range(2):
        xm = VO.synthetic_tile(32, seed=200 + it, batch=4)
        masks = torch.from_numpy(np.stack([VO.mask_generator(rs, 32, 16, 8, 0.5) for _ in range(4)]))
        opt.zero_grad()
        loss, _, _ = mim(xm, masks)
        loss.sum().backward()
        raw = {n: p.grad.detach().clone() for n, p in mim.named_parameters()}
        total = torch.nn.utils.clip_grad_norm_(mim.parameters(), 5.0 if it == 0 else 0.05)   # second step: the clip is active
        opt.step()
        o_loss, o_total, o_grads = TO.train_step(state, tiny_m, xm, masks, clip_grad=5.0 if it == 0 else 0.05)
        assert abs(loss.item() - o_loss.item()) < 1e-6, (loss.item(), o_loss.item())
        assert abs(total.item() - o_total.item()) < 1e-5 * max(1.0, total.item())
        g[f"step{it}/x"], g[f"step{it}/mask"] = xm.numpy(), masks.numpy()
        g[f"step{it}/loss"], g[f"step{it}/grad_norm"] = np.array(loss.item()), np.array(total.item())
        for n, p in mim.named_parameters():
            key = n[len("encoder."):] if n.startswith("encoder.") else n
            assert (p.grad - o_grads[key]).abs().max().item() <= 1e-5 * max(1e-3, p.grad.abs().max().item()), n
            assert (p.detach() - state.params[key]).abs().max().item() <= 2e-6, n
            g[f"step{it}/grad/{key}"] = sample(raw[n])
            g[f"step{it}/gradsum/{key}"] = np.array([float(raw[n].double().sum()), float(raw[n].double().abs().sum())])
            g[f"step{it}/param/{key}"] = sample(p)
    g["dec_w"], g["dec_b"] = dec_w.numpy(), dec_b.numpy()
    for k, v in checksum(sd_m).items():
        g["wsum/" + k] = v
    np.savez_compressed(os.path.join(OUT, "mim_train_tiny.npz"), **g)
    return ["mim_train_tiny.npz: two reference training steps (MIM fwd/bwd, optimizer.py grouping, torch clip_grad_norm_ + AdamW); "
            "oracle/train_oracle.py == reference on loss, grad norm, every gradient and every updated parameter"]


if __name__ == "__main__":
    if "--only" in sys.argv and sys.argv[sys.argv.index("--only") + 1] == "train":
        lines = make_train_golden(_import_reference()[1])
        with open(os.path.join(OUT, "README.md"), "a") as f:
            for r in lines:
                f.write(f"* {r}\n")
        print("\n".join(lines))
    else:
        main()
