"""CPU: the oracle restatement against the golden vectors produced by the reference's own code
(oracle/make_golden.py) and against the installed OpenCV."""
import numpy as np
import pytest
import torch

from conftest import check_weight_sums, load_golden
from oracle import post_oracle as PO
from oracle import vit_oracle as VO

TINY = VO.ViTConfig(embed_dim=128, depth=3, num_heads=2, patch_size=8, img_size=32)


def tiny_sd():
    return VO.randomize_affine(VO.init_state_dict(TINY, seed=7), seed=8)


def test_param_count_known_answers():
    # SSS/output/log_rank0.txt:5570 and :9746
    assert VO.param_count(VO.init_state_dict(VO.ViTConfig(**VO.VIT_SMALL))) == 21670272
    mim = VO.param_count(VO.init_state_dict(VO.ViTConfig(**VO.VIT_SMALL), mim=True)) + 384 * 192 + 192
    assert mim == 21744576


def test_tiny_vit_matches_reference_outputs():
    g = load_golden("tiny_vit.npz")
    sd = tiny_sd()
    check_weight_sums(sd, g)
    for name in ("a", "b", "c"):
        x = torch.from_numpy(g[f"{name}/x"])
        attn = VO.get_last_selfattention(sd, TINY, x)
        feat, attns, qkvs = VO.get_intermediate_feat(sd, TINY, x)
        assert np.abs(attn.numpy() - g[f"{name}/attn"]).max() < 1e-5
        assert np.abs(attns[0].numpy() - g[f"{name}/attn"]).max() < 1e-5
        assert np.abs(feat[0].numpy() - g[f"{name}/feat"]).max() < 1e-4
        assert np.abs(qkvs[0].numpy() - g[f"{name}/qkv"]).max() < 1e-4
        assert np.abs(VO.forward_feats(sd, TINY, x)[:, 0].numpy() - g[f"{name}/cls"]).max() < 1e-4
        rows = VO.cls_attention_rows(sd, TINY, x)
        assert np.allclose(rows.sum(-1).numpy(), 1.0, atol=1e-5)


def test_synthetic_inputs_are_reproducible():
    g = load_golden("tiny_vit.npz")
    for name, (B, S) in {"a": (2, 32), "b": (1, 48), "c": (3, 64)}.items():
        x = VO.synthetic_tile(S, seed=100 + S, batch=B)
        assert np.array_equal(x.numpy(), g[f"{name}/x"])


def test_otsu_and_resize_match_cv2_goldens():
    g = load_golden("cv2_ops.npz")
    for img, t in zip(g["otsu_imgs"], g["otsu_t"]):
        to, mask = PO.otsu_threshold(img)
        assert to == int(t)
        assert np.array_equal(mask, np.where(img > t, 255, 0).astype(np.uint8))
    for s, u in zip(g["resize_in"], g["resize_out"]):
        assert np.abs(PO.resize_linear(s, (56, 56)) - u).max() < 1e-6
        nearest = np.repeat(np.repeat(s, 8, 0), 8, 1)
        assert np.abs(PO.resize_linear(nearest, (7, 7)) - s).max() < 1e-6   # the down-resize is an identity


def test_otsu_and_resize_match_installed_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.RandomState(3)
    for _ in range(20):
        img = np.clip(rng.normal(rng.randint(20, 200), rng.randint(5, 60), (33, 47)), 0, 255).astype(np.uint8)
        t, m = cv2.threshold(img, 0, 255, cv2.THRESH_BINARY + cv2.THRESH_OTSU)
        to, mo = PO.otsu_threshold(img)
        assert int(t) == to and np.array_equal(m, mo)
    s = rng.rand(28, 28).astype(np.float32)
    assert np.abs(cv2.resize(s, (224, 224), interpolation=cv2.INTER_LINEAR) - PO.resize_linear(s, (224, 224))).max() < 1e-6


def test_stitching_matches_reference_loops():
    g = load_golden("stitch.npz")
    for name, (W, S, n) in {"w32s16n4": (32, 16, 4), "w48s16n3": (48, 16, 3), "w24s8n5": (24, 8, 5), "w32s16n1": (32, 16, 1)}.items():
        tiles = list(g[name + "/tiles"])
        out = PO.concat_crops_blend(tiles, S, W)
        assert out.dtype == np.float32 and np.array_equal(out, g[name + "/out"])
        if name + "/img" in g:
            img = g[name + "/img"]
            crops = PO.sliding_window(img, S, W)
            assert len(crops) == n * n
            assert np.array_equal(PO.concat_crops_blend(crops, S, W), g[name + "/gray_stitched"])
        # separable-weights property (size independent): every output pixel is a convex combination
        P = PO.blend_profiles(n, S, W)
        assert np.allclose(P.sum(0), 1.0, atol=1e-12)


def test_threshold_flavours_match_reference():
    g = load_golden("threshold.npz")
    o = PO.threshold_sw(g["img"], g["att"])
    assert np.array_equal(o[0], g["sw_th"]) and np.array_equal(o[1], g["sw_th2"]) and np.array_equal(o[2], g["sw_th3"])
    u = PO.threshold_utils(g["img"], g["att"])
    assert np.array_equal(u[0], g["ut_th"]) and np.array_equal(u[1], g["ut_th2"]) and np.array_equal(u[2], g["ut_th3"])


def test_vits8_tile_post_chain_matches_reference():
    g = load_golden("vits8_tile.npz")
    x = VO.synthetic_tile(224, seed=int(g["x_seed"]), batch=1)
    rows = g["cls_rows"][0]
    att = PO.tile_attention_map(rows, 224, 8)
    assert np.abs(att - g["att_map"]).max() <= 1e-6 * np.abs(g["att_map"]).max()
    th, th2, th3, _, _ = PO.eval_tile(rows, x[0, 0].numpy(), 8)
    for a, b in ((th, g["th"]), (th2, g["th2"]), (th3, g["th3"])):
        assert (a == b).mean() >= 0.9999


def test_vitb8_oracle_matches_reference_golden():
    """ViT-B/8 (BASELINE configs[2]): the oracle's functional forward against the reference's own CLS rows and masks."""
    g = load_golden("vitb8_tile.npz")
    cfg = VO.ViTConfig(**VO.VIT_BASE)
    sd = VO.randomize_affine(VO.init_state_dict(cfg, seed=0), seed=1, scale=0.02)
    check_weight_sums(sd, g)
    x = VO.synthetic_tile(224, seed=int(g["x_seed"]), batch=1)
    rows = VO.cls_attention_rows(sd, cfg, x).numpy()
    assert rows.shape == (1, 12, 785)
    assert (np.abs(rows - g["cls_rows"]) / np.abs(g["cls_rows"])).max() <= 1e-5
    th, th2, th3, _, _ = PO.eval_tile(rows[0], x[0, 0].numpy(), 8)
    for a, b in ((th, g["th"]), (th2, g["th2"]), (th3, g["th3"])):
        assert (a == b).mean() >= 0.9999


def test_mask_generator_and_mim_loss():
    g = load_golden("mim_tiny.npz")
    m = VO.mask_generator(np.random.RandomState(0), 224, 16, 8, 0.5)
    assert np.array_equal(m, g["mask224_seed0"]) and m.sum() == 392
    cfg_init = VO.ViTConfig(embed_dim=128, depth=2, num_heads=2, patch_size=8, img_size=224)
    cfg = VO.ViTConfig(embed_dim=128, depth=2, num_heads=2, patch_size=8, img_size=32)
    sd = VO.randomize_affine(VO.init_state_dict(cfg_init, seed=11, mim=True), seed=12)
    check_weight_sums(sd, g)
    loss, x_rec, _ = VO.mim_forward(sd, cfg, torch.from_numpy(g["dec_w"]), torch.from_numpy(g["dec_b"]),
                                    torch.from_numpy(g["x"]), torch.from_numpy(g["mask"]))
    assert abs(loss.item() - float(g["loss"])) < 1e-6
    assert np.abs(x_rec.detach().numpy() - g["x_rec"]).max() < 1e-5


def _sample(t):
    f = t.detach().reshape(-1)
    return (f if t.dim() <= 1 or f.numel() <= 4096 else f[::7]).numpy()


def test_train_step_oracle_matches_reference_golden():
    """oracle/train_oracle.py (autograd over the functional forward + restated clip_grad_norm_ / AdamW / param grouping)
    against two iterations of the reference's own training step (tests/golden/mim_train_tiny.npz)."""
    from oracle import train_oracle as TO
    g = load_golden("mim_train_tiny.npz")
    cfg_init = VO.ViTConfig(embed_dim=128, depth=2, num_heads=2, patch_size=8, img_size=224)
    cfg = VO.ViTConfig(embed_dim=128, depth=2, num_heads=2, patch_size=8, img_size=32)
    sd = VO.randomize_affine(VO.init_state_dict(cfg_init, seed=11, mim=True), seed=12)
    check_weight_sums(sd, g)
    params = dict(sd)
    params["decoder.0.weight"], params["decoder.0.bias"] = torch.from_numpy(g["dec_w"]), torch.from_numpy(g["dec_b"])
    state = TO.TrainState(params)
    for it, clip in ((0, 5.0), (1, 0.05)):
        x, mask = torch.from_numpy(g[f"step{it}/x"]), torch.from_numpy(g[f"step{it}/mask"])
        loss, raw = TO.mim_loss_and_grads(state.params, cfg, x, mask)
        for k, gr in raw.items():
            ref = g[f"step{it}/grad/{k}"]
            assert np.abs(_sample(gr) - ref).max() <= 1e-5 * max(1e-3, np.abs(ref).max()), k
            assert np.allclose([float(gr.double().sum()), float(gr.double().abs().sum())], g[f"step{it}/gradsum/{k}"], rtol=1e-4, atol=1e-6), k
        loss2, total, _ = TO.train_step(state, cfg, x, mask, clip_grad=clip)
        assert abs(loss.item() - float(g[f"step{it}/loss"])) < 1e-6 and abs(loss2.item() - loss.item()) < 1e-7
        assert abs(total.item() - float(g[f"step{it}/grad_norm"])) <= 1e-5 * max(1.0, float(g[f"step{it}/grad_norm"]))
        for k, p in state.params.items():
            assert np.abs(_sample(p) - g[f"step{it}/param/{k}"]).max() <= 2e-6, k
    assert float(g["step1/grad_norm"]) > 0.05      # the second step's clip (max_norm 0.05) was active


def test_adamw_restatement_matches_torch_optim():
    from oracle import train_oracle as TO
    gen = torch.Generator().manual_seed(0)
    p0 = torch.randn(1000, generator=gen)
    p_ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([p_ref], lr=5e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.05)
    p, m, v = p0.clone(), torch.zeros(1000), torch.zeros(1000)
    for step in range(1, 6):
        gr = torch.randn(1000, generator=gen) * 0.1
        p_ref.grad = gr.clone()
        opt.step()
        p, m, v = TO.adamw_update(p, gr, m, v, step, 5e-4, 0.9, 0.999, 1e-8, 0.05)
        assert (p - p_ref.detach()).abs().max().item() < 1e-6


def test_lr_scheduler_and_param_groups_mirror():
    """Host logic of the mirrors (no GPU): cosine schedule == the oracle's timm formula; parameter grouping of
    optimizer.get_pretrain_param_groups == SSS/optimizer.py:14-33 on the MIM module."""
    from functools import partial
    from types import SimpleNamespace as NS
    import vitocm_b200 as vob
    from oracle import train_oracle as TO
    enc = vob.VisionTransformerForSimMIM(patch_size=8, embed_dim=128, depth=2, num_heads=2, mlp_ratio=4, img_size=[32], qkv_bias=True,
                                         norm_layer=partial(torch.nn.LayerNorm, eps=1e-6))
    mim = vob.MIM(encoder=enc, encoder_stride=8)
    assert mim.no_weight_decay() == set() or mim.no_weight_decay() == {}
    groups = vob.optimizer.get_pretrain_param_groups(mim, None, mim.no_weight_decay(), mim.no_weight_decay_keywords())
    decay_ids = {id(p) for p in groups[0]["params"]}
    for n, p in mim.named_parameters():
        key = n[len("encoder."):] if n.startswith("encoder.") else n
        assert (id(p) in decay_ids) == TO.has_weight_decay(key, p.shape), n
    assert groups[1]["weight_decay"] == 0.0
    opt = torch.optim.SGD([{"params": groups[0]["params"]}, {"params": groups[1]["params"]}], lr=5e-4)
    cfg = NS(TRAIN=NS(EPOCHS=10, WARMUP_EPOCHS=2, MIN_LR=5e-6, WARMUP_LR=5e-7, LR_SCHEDULER=NS(NAME="cosine", DECAY_EPOCHS=30, MULTISTEPS=[])))
    sched = vob.lr_scheduler.build_scheduler(cfg, opt, n_iter_per_epoch=7)
    assert abs(opt.param_groups[0]["lr"] - 5e-7) < 1e-12          # timm initialises the groups to warmup_lr_init
    for t in (0, 1, 13, 14, 15, 40, 69, 70, 100):
        sched.step_update(t)
        want = TO.cosine_lr(t, 5e-4, 70, 5e-6, 14, 5e-7)
        assert all(abs(gp["lr"] - want) < 1e-12 for gp in opt.param_groups), t


def test_package_synthetic_inputs_equal_the_oracle_recipe():
    """bench.py's GPU arm draws its inputs from vitocm_b200.synthetic (it may not touch oracle/); the recipe must not drift."""
    from vitocm_b200 import synthetic as SY
    assert torch.equal(SY.synthetic_tile(64, seed=5, batch=2), VO.synthetic_tile(64, seed=5, batch=2))
    assert np.array_equal(SY.synthetic_mosaic_u8(160, seed=6), VO.synthetic_mosaic_u8(160, seed=6))
    a = SY.random_masks(np.random.RandomState(3), 4, 32, 16, 8, 0.5).numpy()
    rs = np.random.RandomState(3)
    b = np.stack([VO.mask_generator(rs, 32, 16, 8, 0.5) for _ in range(4)])
    assert np.array_equal(a, b)


def test_crop_and_overlap_variants_match_reference():
    """concat_crops_overlap (SSS/utils.py:319-347), utils.sliding_window (:349-362) and the
    `--crop 4|16` evaluation path (SSS/eval.py:145-173) against outputs of the reference's own functions."""
    g = load_golden("variants.npz")
    for name, st in {"w16s2n3": 2, "w16s5n4": 5, "w12s3n2": 3, "w10s2n1": 2, "w24s4n5": 4}.items():
        for kind in ("f32", "u8", "rgb"):
            tiles = list(g[f"overlap/{name}/{kind}/tiles"])
            out = PO.concat_crops_overlap(tiles, st)
            want = g[f"overlap/{name}/{kind}/out"]
            assert out.dtype == want.dtype and np.array_equal(out, want), (name, kind)
    crops = PO.sliding_window_utils(g["sw/img"], 24, 10)
    assert np.array_equal(np.stack(crops), g["sw/crops"])
    for name in ("crop4", "crop16"):
        images, rows = g[f"{name}/images"], g[f"{name}/cls_rows"]
        for b in range(images.shape[0]):
            att, th = PO.eval_cropped(rows[b], images[b, :, 0], 8)
            assert np.abs(att - g[f"{name}/attention"][b]).max() < 1e-6
            for k in range(3):
                assert np.array_equal(th[k], g[f"{name}/masks"][b, k])
