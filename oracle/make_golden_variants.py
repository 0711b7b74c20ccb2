"""Generate tests/golden/variants.npz (crop / overlap stitching variants) by running
the REFERENCE's own code from /root/reference, and check the oracle restatement against it.

Run in the build container only:  PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_golden_variants

  * utils.py:304-362 (`concat_crops`, `concat_crops_overlap`, `sliding_window`) : function bodies exec'd from the
    source lines (the module itself needs matplotlib / scikit-image);
  * eval.py:145-173 (`--crop 4`) : the loop body cannot be imported (argparse / wandb at module level), so it is
    driven here call by call with the reference's own pieces -- its ViT (`get_intermediate_feat`),
    `utils.compute_attention`, numpy mean, `utils.concat_crops`, the two `cv2.resize` calls, torchvision's
    `ToPILImage` + `convert("L")`, and `utils.threshold`.
Only OUTPUTS are stored.
"""
from __future__ import annotations

import os
from functools import partial

import numpy as np
import torch

from oracle.make_golden import OUT, REF, _exec_lines, _import_reference, _ref_post_namespaces


def main():
    import cv2
    from PIL import Image
    from oracle import post_oracle as PO
    from oracle import vit_oracle as VO

    vits, _ = _import_reference()
    ns_utils, _ = _ref_post_namespaces()
    ns_var = _exec_lines(os.path.join(REF, "utils.py"), [(304, 362)], dict(np=np))
    rng = np.random.RandomState(77)
    g = {}

    # ---- concat_crops_overlap: float32 and uint8 (gray + RGB), overlaps below and above half a window
    for name, (W, st, n) in {"w16s2n3": (16, 2, 3), "w16s5n4": (16, 5, 4), "w12s3n2": (12, 3, 2), "w10s2n1": (10, 2, 1),
                             "w24s4n5": (24, 4, 5)}.items():
        tf = [((rng.rand(W, W) - 0.3) * 300).astype(np.float32) for _ in range(n * n)]
        tu = [rng.randint(0, 256, (W, W)).astype(np.uint8) for _ in range(n * n)]
        tc = [rng.randint(0, 256, (W, W, 3)).astype(np.uint8) for _ in range(n * n)]
        for kind, tiles in (("f32", tf), ("u8", tu), ("rgb", tc)):
            ref = ns_var["concat_crops_overlap"](tiles, st)
            orc = PO.concat_crops_overlap(tiles, st)
            assert ref.dtype == orc.dtype and np.array_equal(ref, orc), (name, kind)
            g[f"overlap/{name}/{kind}/tiles"] = np.stack(tiles)
            g[f"overlap/{name}/{kind}/out"] = ref
    # ---- utils.sliding_window (window, stride order) on a ragged RGB image
    img = rng.randint(0, 256, (70, 90, 3)).astype(np.uint8)
    ref = ns_var["sliding_window"](Image.fromarray(img), 24, 10)
    orc = PO.sliding_window_utils(img, 24, 10)
    assert len(ref) == len(orc) and all(np.array_equal(a, b) for a, b in zip(ref, orc))
    g["sw/img"] = img
    g["sw/crops"] = np.stack(ref)

    # ---- eval.py `--crop 4` / `--crop 16` on a tiny ViT
    import torchvision.transforms as T
    tiny = VO.ViTConfig(embed_dim=128, depth=3, num_heads=2, patch_size=8, img_size=32)
    sd_t = VO.randomize_affine(VO.init_state_dict(tiny, seed=7), seed=8)
    model = vits.VisionTransformer(img_size=[32], patch_size=8, embed_dim=128, depth=3, num_heads=2, mlp_ratio=4, qkv_bias=True,
                                   norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), num_classes=0)
    model.load_state_dict(sd_t, strict=True)
    model.eval()
    to_pil = T.ToPILImage()
    for name, (cr, s, B) in {"crop4": (2, 48, 2), "crop16": (4, 32, 1)}.items():
        S = cr * s
        full = VO.synthetic_tile(S, seed=300 + S, batch=B)                       # [B, 3, S, S], R = G = B
        images = torch.stack([torch.stack([full[b, :, i * s:(i + 1) * s, j * s:(j + 1) * s] for i in range(cr) for j in range(cr)])
                              for b in range(B)])                                # [B, cr*cr, 3, s, s] (data.py:115-121)
        atts, ths, rows_all = [], [], []
        for i in range(B):
            average_crops, rows_i = [], []
            for j in range(images.shape[1]):
                crop = images[i][j].unsqueeze(0)
                with torch.no_grad():
                    feat, attentions, qkv = model.get_intermediate_feat(crop, n=1)
                w_f, h_f = crop.shape[-2] // 8, crop.shape[-1] // 8
                resp, nh = ns_utils["compute_attention"](attentions, 0, w_f, h_f, 8)
                average_crops.append(np.mean(resp, axis=0))
                rows_i.append(attentions[0][0, :, 0, :].numpy())
            average = ns_var["concat_crops"](average_crops)
            img_t = torch.tensor(ns_var["concat_crops"](images[i, :, 0, :, :]))
            temp = torch.zeros([1, 3, S, S], dtype=torch.float32)
            temp[0][0] = img_t
            temp[0][1] = img_t
            temp[0][2] = img_t
            average = cv2.resize(average, (average.shape[1] // 8, average.shape[0] // 8))
            average = cv2.resize(average, (temp.shape[-1], temp.shape[-1]), interpolation=cv2.INTER_LINEAR)
            out = ns_utils["threshold"](to_pil(temp.squeeze(0)).convert("L"), average, save=False)
            rows_i = np.stack(rows_i)
            o_att, o_th = PO.eval_cropped(rows_i, images[i, :, 0].numpy(), 8)
            assert np.abs(o_att - average).max() < 1e-6, (name, np.abs(o_att - average).max())
            agree = [float((a == b).mean()) for a, b in zip(out, o_th)]
            assert min(agree) == 1.0, (name, agree)
            atts.append(average)
            ths.append(np.stack(out))
            rows_all.append(rows_i)
        g[f"{name}/images"] = images.numpy()
        g[f"{name}/cls_rows"] = np.stack(rows_all)
        g[f"{name}/attention"] = np.stack(atts)
        g[f"{name}/masks"] = np.stack(ths)
    np.savez_compressed(os.path.join(OUT, "variants.npz"), **g)
    line = ("variants.npz: concat_crops_overlap (f32 / u8 / RGB, 5 geometries), utils.sliding_window and the eval.py "
            "`--crop 4|16` path: oracle restatement bit-equal to the reference's functions (attention within 1e-6)")
    print(line)


if __name__ == "__main__":
    main()
