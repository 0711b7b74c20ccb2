"""GPU (B200): post-processing / stitching kernels against the golden vectors and the oracle.
Integer and byte outputs must be bit-exact; float32 stitched maps must be bit-exact too."""
import numpy as np
import pytest
import torch

import vitocm_b200 as vob
from conftest import load_golden
from gpu_util import build_model
from oracle import post_oracle as PO
from oracle import vit_oracle as VO
from vitocm_b200 import sw_processing as sw
from vitocm_b200 import utils as vu
from vitocm_b200._lib import check, cur_stream, ptr

pytestmark = pytest.mark.gpu


def test_otsu_kernel_matches_cv2_goldens():
    g = load_golden("cv2_ops.npz")
    hists = np.stack([np.bincount(img.ravel(), minlength=256) for img in g["otsu_imgs"]]).astype(np.int64)
    d = torch.from_numpy(hists).cuda()
    thr = torch.empty(len(hists), dtype=torch.int32, device="cuda")
    check(vob._lib.load_library().vitocm_otsu(ptr(d), len(hists), ptr(thr), cur_stream()))
    assert np.array_equal(thr.cpu().numpy(), g["otsu_t"])
    # degenerate histograms
    deg = torch.zeros(3, 256, dtype=torch.int64, device="cuda")
    deg[1, 17] = 100
    deg[2, 0] = 5
    deg[2, 255] = 5
    t = torch.empty(3, dtype=torch.int32, device="cuda")
    check(vob._lib.load_library().vitocm_otsu(ptr(deg), 3, ptr(t), cur_stream()))
    exp = [PO.otsu_from_hist(h) for h in deg.cpu().numpy()]
    assert t.cpu().tolist() == exp


def test_threshold_flavours_bit_exact():
    g = load_golden("threshold.npz")
    th, th2, th3 = sw.threshold(g["img"], g["att"], save=False)
    assert np.array_equal(th, g["sw_th"]) and np.array_equal(th2, g["sw_th2"]) and np.array_equal(th3, g["sw_th3"])
    th, th2, th3 = vu.threshold(g["img"], g["att"], save=False)
    assert np.array_equal(th, g["ut_th"]) and np.array_equal(th2, g["ut_th2"]) and np.array_equal(th3, g["ut_th3"])
    # flat attention: min_max_normalize returns its input
    flat = np.full((16, 16), 0.5, np.float32)
    img = np.random.RandomState(0).randint(0, 256, (16, 16)).astype(np.uint8)
    for fn, orc in ((sw.threshold, PO.threshold_sw), (vu.threshold, PO.threshold_utils)):
        got = fn(img, flat, save=False)
        exp = orc(img, flat)
        assert all(np.array_equal(a, b) for a, b in zip(got, exp[:3]))


def test_threshold_saves_the_reference_files_bit_exact(tmp_path):
    """save=True writes the reference's five PNGs (SSS/utils.py:102-114, SSS/sw_processing.py:68-80); the masks,
    the weighted image `result` and the normalised attention read back from disk equal the oracle's arrays."""
    from PIL import Image
    g = load_golden("threshold.npz")
    for fn, orc, sub in ((vu.threshold, PO.threshold_utils, "ut"), (sw.threshold, PO.threshold_sw, "sw")):
        out = tmp_path / sub
        fn(g["img"], g["att"], output_directory=str(out), save=True, name="case")
        th, th2, th3, result, att_u8 = orc(g["img"], g["att"])
        files = {"case/OTSU_th_average.png": th, "OTSU_th_original.png": th2, "weighted_iamge_attention.png": result,
                 "heatmap_otsu_attention.png": th3, "temp.png": att_u8}
        for rel, want in files.items():
            assert np.array_equal(np.array(Image.open(out / rel)), want), (sub, rel)


@pytest.mark.parametrize("name,W,S,n", [("w32s16n4", 32, 16, 4), ("w48s16n3", 48, 16, 3), ("w24s8n5", 24, 8, 5), ("w32s16n1", 32, 16, 1)])
def test_concat_crops_and_sliding_window_bit_exact(name, W, S, n):
    g = load_golden("stitch.npz")
    out = sw.concat_crops(list(g[name + "/tiles"]), S, W)
    assert out.dtype == np.float32 and np.array_equal(out, g[name + "/out"])
    if name + "/img" in g:
        img = g[name + "/img"]
        rgb = np.stack([img] * 3, -1)
        crops = sw.sliding_window(rgb, S, W)
        ocrops = PO.sliding_window(rgb, S, W)
        assert len(crops) == len(ocrops) == n * n and all(np.array_equal(a, b) for a, b in zip(crops, ocrops))
        st = sw.concat_crops(crops, S, W)
        assert np.array_equal(st[..., 0], g[name + "/gray_stitched"])
        # device-resident variant used by MosaicSegmenter
        d = torch.from_numpy(img).cuda()
        E = (n - 1) * S + W
        gray = torch.zeros(E, E, dtype=torch.uint8, device="cuda")
        wtab = sw._wtab(W, S, d.device)
        check(vob._lib.load_library().vitocm_stitch_gray(ptr(d), img.shape[0], img.shape[1], d.stride(0), n, W, S, ptr(wtab), 0, E,
                                                         ptr(gray), cur_stream()))
        assert np.array_equal(gray.cpu().numpy(), g[name + "/gray_stitched"])


def test_sliding_window_ragged_image_zero_pads():
    rng = np.random.RandomState(1)
    img = rng.randint(0, 256, (100, 100, 3)).astype(np.uint8)       # 100 is not a multiple of the stride
    got = sw.sliding_window(img, 16, 48)
    exp = PO.sliding_window(img, 16, 48)
    assert len(got) == len(exp) and all(np.array_equal(a, b) for a, b in zip(got, exp))
    assert sw.sliding_window(np.zeros((20, 20), np.uint8), 16, 48) == []


def test_head_mean_and_tile_threshold_vs_oracle():
    rng = np.random.RandomState(2)
    T, H, S, p = 5, 6, 64, 8
    n = (S // p) ** 2
    rows = rng.rand(T, H, n + 1).astype(np.float32)
    rows /= rows.sum(-1, keepdims=True)
    x = VO.synthetic_tile(S, seed=11, batch=T)
    d_rows = torch.from_numpy(rows).cuda()
    low = vob.head_mean_maps(d_rows).cpu().numpy()
    for t in range(T):
        a, _ = PO.compute_attention_from_rows(rows[t], S // p, S // p, p)
        assert np.array_equal(low[t].reshape(S // p, S // p), np.mean(a, axis=0)[::p, ::p])
    low255 = vob.head_mean_maps(d_rows, per_tile_minmax255=True).cpu().numpy()
    for t in range(T):
        assert np.array_equal(PO.resize_linear(low255[t].reshape(8, 8), (S, S)), PO.sw_tile_map(rows[t], S, p))
    masks = torch.empty(T, 3, S, S, dtype=torch.uint8, device="cuda")
    thr = torch.empty(T, 3, dtype=torch.int32, device="cuda")
    att = torch.empty(T, S, S, device="cuda")
    lowd = torch.from_numpy(low).cuda()
    check(vob._lib.load_library().vitocm_tile_threshold(ptr(lowd), ptr(x.cuda()), T, 3, S, S // p, S // p, ptr(masks), ptr(thr),
                                                        ptr(att), None, None, cur_stream()))
    for t in range(T):
        assert np.array_equal(att[t].cpu().numpy(), PO.tile_attention_map(rows[t], S, p))
        th, th2, th3, _, _ = PO.eval_tile(rows[t], x[t, 0].numpy(), p)
        got = masks[t].cpu().numpy()
        assert np.array_equal(got[0], th) and np.array_equal(got[1], th2) and np.array_equal(got[2], th3)


@pytest.mark.parametrize("W,S,size", [(32, 16, 112), (48, 16, 128), (32, 16, 100)])
def test_mosaic_pipeline_vs_oracle(W, S, size):
    """Whole sliding-window pipeline on a small mosaic with a tiny ViT: the ViT stage is compared
    within tolerance, everything after it (fed with the GPU's own CLS rows) bit-exactly."""
    tiny = VO.ViTConfig(embed_dim=128, depth=2, num_heads=2, patch_size=8, img_size=W)
    sd = VO.randomize_affine(VO.init_state_dict(tiny, seed=21), seed=22)
    m = build_model(tiny, sd, "fp32", chunk_tiles=5)
    mosaic = VO.synthetic_mosaic_u8(size, seed=77)
    seg = vob.MosaicSegmenter(m, window=W, stride=S, tile_batch=7)
    out = seg.segment(torch.from_numpy(mosaic).cuda(), want=("th", "th2", "th3"))
    n = vob.grid_size(size, S)
    assert out["grid"] == n and out["extent"] == (n - 1) * S + W
    # ViT stage vs oracle
    crops = PO.sliding_window(mosaic, S, W)
    xs = torch.from_numpy(np.stack(crops)).float().div(255.0).unsqueeze(1).expand(-1, 3, -1, -1).contiguous()
    rows_ref = VO.cls_attention_rows(sd, tiny, xs).numpy()
    rows_gpu = m.cls_attention_rows(xs[:, :1].contiguous().cuda()).cpu().numpy()     # gray fast path, as MosaicSegmenter runs it
    assert float((np.abs(rows_gpu - rows_ref) / rows_ref).max()) <= 1e-3
    rows_rgb = m.cls_attention_rows(xs.cuda()).cpu().numpy()                          # the 3-channel path: same up to summation order
    assert float((np.abs(rows_gpu - rows_rgb) / rows_rgb).max()) <= 2e-5
    # post stage, exact, from the GPU's rows
    stitched, (th, th2, th3, _, _), gray = PO.mosaic_segment(rows_gpu, mosaic, S, W, 8)
    assert np.array_equal(seg.stitched_map(out["lowres"]).cpu().numpy(), stitched)
    assert np.array_equal(out["th"].cpu().numpy(), th)
    assert np.array_equal(out["th2"].cpu().numpy(), th2)
    assert np.array_equal(out["th3"].cpu().numpy(), th3)
    # and end to end against the pure oracle
    _, (th_o, _, th3_o, _, _), _ = PO.mosaic_segment(rows_ref, mosaic, S, W, 8)
    assert float((out["th"].cpu().numpy() == th_o).mean()) >= 0.999
    assert float((out["th3"].cpu().numpy() == th3_o).mean()) >= 0.999


def test_pgt_pseudo_masks_match_the_per_image_reference_loop():
    """the pseudo-mask generation of SSS/PGT.py:55-91 for a whole batch on the device, against the
    oracle's per-image restatement (compute_attention -> head mean -> resize pair -> utils.threshold); all heads, and the
    ``rand`` branch with the reference's numpy draw order."""
    tiny = VO.ViTConfig(embed_dim=128, depth=2, num_heads=2, patch_size=8, img_size=64)
    sd = VO.randomize_affine(VO.init_state_dict(tiny, seed=21), seed=22)
    m = build_model(tiny, sd, "fp32", chunk_tiles=4)
    x = VO.synthetic_tile(64, seed=91, batch=5)
    rows = VO.cls_attention_rows(sd, tiny, x).numpy()
    y = vob.pgt.pseudo_masks(m, x.cuda())
    assert tuple(y.shape) == (5, 1, 64, 64) and set(np.unique(y.cpu().numpy()).tolist()) <= {0.0, 1.0}
    for b in range(5):
        th, _, _, _, _ = PO.eval_tile(rows[b], x[b, 0].numpy(), 8)
        assert float(((y[b, 0].cpu().numpy() * 255).astype(np.uint8) == th).mean()) >= 0.999
    # the rand branch draws 1..6 heads (SSS/PGT.py:67-69 assumes the 6 heads of ViT-S): a one-block ViT-S-wide model
    six = VO.ViTConfig(embed_dim=384, depth=1, num_heads=6, patch_size=8, img_size=64)
    sd6 = VO.randomize_affine(VO.init_state_dict(six, seed=23), seed=24)
    m6 = build_model(six, sd6, "fp32", chunk_tiles=4)
    rows6 = VO.cls_attention_rows(sd6, six, x).numpy()
    np.random.seed(5)
    w = vob.pgt.select_heads(5, 6)
    np.random.seed(5)
    y2 = vob.pgt.pseudo_masks(m6, x.cuda(), rand=True)
    assert (w.sum(1) - 1).abs().max().item() < 1e-6
    for b in range(5):
        sel = np.nonzero(w[b].numpy())[0]
        th, _, _, _, _ = PO.eval_tile(rows6[b][sel], x[b, 0].numpy(), 8)
        assert float(((y2[b, 0].cpu().numpy() * 255).astype(np.uint8) == th).mean()) >= 0.999


def test_concat_crops_overlap_and_utils_sliding_window_bit_exact():
    """SSS/utils.py:319-347 / :349-362 on the device against the reference's golden outputs
    (float32, uint8 gray, uint8 RGB; overlaps below and above half a window; n = 1)."""
    g = load_golden("variants.npz")
    for name, st in {"w16s2n3": 2, "w16s5n4": 5, "w12s3n2": 3, "w10s2n1": 2, "w24s4n5": 4}.items():
        for kind in ("f32", "u8", "rgb"):
            tiles = list(g[f"overlap/{name}/{kind}/tiles"])
            out = vu.concat_crops_overlap(tiles, st)
            want = g[f"overlap/{name}/{kind}/out"]
            assert out.dtype == want.dtype and np.array_equal(out, want), (name, kind)
    crops = vu.sliding_window(g["sw/img"], 24, 10)
    assert np.array_equal(np.stack(crops), g["sw/crops"])
    # a larger random case against the oracle (all three dtypes in one geometry the goldens do not hold)
    rng = np.random.RandomState(5)
    tiles = [((rng.rand(40, 40) - 0.5) * 1000).astype(np.float32) for _ in range(36)]
    assert np.array_equal(vu.concat_crops_overlap(tiles, 7), PO.concat_crops_overlap(tiles, 7))
    tiles = [rng.randint(0, 256, (40, 40, 3)).astype(np.uint8) for _ in range(36)]
    assert np.array_equal(vu.concat_crops_overlap(tiles, 13), PO.concat_crops_overlap(tiles, 13))
    with pytest.raises(vob.VitocmError):
        vu.concat_crops_overlap(tiles, 20)          # 2 * stride must stay below the window


@pytest.mark.parametrize("name", ["crop4", "crop16"])
def test_cropped_evaluation_path_matches_reference_goldens(name):
    """`--crop 4|16` (SSS/eval.py:145-173, SSS/data.py:85-125) batched on the device against the
    reference's own run: CLS rows within 1e-3, masks >= 99.9 %; and bit-exact against the oracle fed the GPU's rows."""
    g = load_golden("variants.npz")
    tiny = VO.ViTConfig(embed_dim=128, depth=3, num_heads=2, patch_size=8, img_size=32)
    sd = VO.randomize_affine(VO.init_state_dict(tiny, seed=7), seed=8)
    m = build_model(tiny, sd, "fp32", chunk_tiles=5)
    images = torch.from_numpy(g[f"{name}/images"])
    out = vu.cropped_attention_masks(m, images.cuda(), return_attention=True)
    B, ncrop = images.shape[:2]
    rows_gpu = out["cls_rows"].cpu().numpy().reshape(B, ncrop, 2, -1)
    rows_ref = g[f"{name}/cls_rows"]
    assert float((np.abs(rows_gpu - rows_ref) / rows_ref).max()) <= 1e-3
    masks = out["masks"].cpu().numpy()
    att = out["attention"].cpu().numpy()
    for b in range(B):
        o_att, o_th = PO.eval_cropped(rows_gpu[b], images[b, :, 0].numpy(), 8)
        assert np.array_equal(att[b], o_att)
        for k in range(3):
            assert np.array_equal(masks[b, k], o_th[k])
            assert float((masks[b, k] == g[f"{name}/masks"][b, k]).mean()) >= 0.999
        assert np.abs(att[b] - g[f"{name}/attention"][b]).max() <= 1e-3 * np.abs(g[f"{name}/attention"][b]).max()


def test_direct_mosaic_ingest_equals_materialised_crops_bit_for_bit():
    """the patch-embedding producer cuts its tiles out of the uint8 mosaic (v / 255 by a correctly rounded
    reciprocal product, zero beyond the mosaic) -- same CLS rows, bit for bit, as fp32 crops cut by vitocm_extract_tiles, on an
    aligned mosaic, a ragged one (windows hanging over the edge, odd pitch) and a row-strided view; all 256 byte values occur."""
    tiny = VO.ViTConfig(embed_dim=128, depth=2, num_heads=2, patch_size=8, img_size=32)
    sd = VO.randomize_affine(VO.init_state_dict(tiny, seed=21), seed=22)
    m = build_model(tiny, sd, "bf16", chunk_tiles=5)
    rng = np.random.RandomState(3)
    lib = vob._lib.load_library()
    for size, W, S, view in ((112, 32, 16, False), (101, 32, 16, False), (67, 24, 8, False), (96, 32, 16, True)):
        base = rng.randint(0, 256, (size, size + (13 if view else 0))).astype(np.uint8)
        base.flat[:256] = np.arange(256, dtype=np.uint8)
        full = torch.from_numpy(base).cuda()
        mosaic = full[:, 5:5 + size] if view else full            # a view: column offset 5 breaks the 8-byte alignment, pitch != width
        n = vob.grid_size(size, S)
        T = n * n
        x = torch.empty(T, 1, W, W, device="cuda")
        check(lib.vitocm_extract_tiles(mosaic.data_ptr(), size, size, mosaic.stride(0), n, W, S, 0, T, 1, ptr(x), cur_stream()))
        want = m.cls_attention_rows(x)
        got = m.cls_attention_rows_mosaic(mosaic, n, W, S, 0, T)
        assert torch.equal(got, want), (size, W, S, view)
        part = m.cls_attention_rows_mosaic(mosaic, n, W, S, 3, T - 4)         # a sub-range of the windows
        assert torch.equal(part, want[3:T - 1])
    seg_a = vob.MosaicSegmenter(m, window=32, stride=16, tile_batch=7, ingest="direct")
    seg_b = vob.MosaicSegmenter(m, window=32, stride=16, tile_batch=7, ingest="crops")
    mos = torch.from_numpy(VO.synthetic_mosaic_u8(100, seed=5)).cuda()
    a, b = seg_a.segment(mos), seg_b.segment(mos)
    assert torch.equal(a["lowres"], b["lowres"]) and torch.equal(a["th"], b["th"]) and torch.equal(a["th3"], b["th3"])


@pytest.mark.parametrize("threshold", [0.6, 0.9, 0.1])
def test_cumulative_mass_threshold_matches_the_upstream_restatement(threshold):
    """`--threshold` (SSS/eval.py:33-34): keep the patches holding the top `threshold` of each head's attention mass.  Oracle =
    the PyTorch restatement of upstream DINO's visualize_attention.py (parity unpinned by the reference).  Sorting and the mask
    scatter are exact; the cumulative sum is a parallel scan here and sequential in torch, so an element whose cumulative mass
    lies within 1e-5 of 1 - threshold may fall on either side -- those are excluded, everything else must agree exactly."""
    rng = np.random.RandomState(7)
    T, H, N = 5, 6, 785
    logits = rng.randn(T, H, N).astype(np.float32) * 1.5
    rows = np.exp(logits) / np.exp(logits).sum(-1, keepdims=True)
    rows[1, 2, 5:9] = rows[1, 2, 5]                      # ties
    d = torch.from_numpy(rows).cuda()
    mask, up = vob.utils.cummass_threshold(d, threshold, 28, 28, 8)
    mask, up = mask.cpu().numpy(), up.cpu().numpy()
    assert mask.shape == (T, H, N - 1) and up.shape == (T, H, 224, 224)
    for t in range(T):
        o_up, o_low, (cumval, idx) = PO.cummass_threshold(rows[t], threshold, 28, 28, 8)
        near = np.zeros((H, N - 1), dtype=bool)
        for h in range(H):
            near[h, idx[h][np.abs(cumval[h] - np.float32(1 - threshold)) <= 1e-5]] = True
        assert ((mask[t] != 0) == o_low)[~near].all()
        assert near.sum() <= 4 * H
        # kept mass is at least `threshold` of the head's total, and dropping the smallest kept patch would go below it
        for h in range(H):
            kept = rows[t, h, 1:][mask[t, h] != 0].sum() / rows[t, h, 1:].sum()
            assert kept >= threshold - 1e-4
        # nearest x 8 upsampling of the device's own low-res mask
        assert np.array_equal(up[t], np.repeat(np.repeat(mask[t].reshape(H, 28, 28), 8, 1), 8, 2).astype(np.float32))
