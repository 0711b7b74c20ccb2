"""Generate tests/golden/vitb8_tile.npz: the REFERENCE's own ViT-B/8 (SSS/dino/vision_transformer.py:275-279, imported
read-only from /root/reference) on one seeded synthetic 224 x 224 tile -- last-layer CLS attention rows through
get_intermediate_feat (the call SSS/sw_processing.py:239 makes), the eval-flavour masks of the reference's own
compute_attention / threshold (exec'd from SSS/utils.py), and weight checksums.  BASELINE.json configs[2] runs this model.

Run in the build container only (the GPU box has no /root/reference):
    PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_golden_vitb
Nothing from the reference is copied into this repository: only its OUTPUTS are stored.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from oracle.make_golden import OUT, _import_reference, _ref_post_namespaces, checksum


def main():
    import cv2
    import torchvision.transforms as T
    from oracle import post_oracle as PO
    from oracle import vit_oracle as VO

    torch.set_num_threads(max(1, os.cpu_count() or 1))
    vits, _ = _import_reference()
    ns_utils, _ = _ref_post_namespaces()
    cfg = VO.ViTConfig(**VO.VIT_BASE)
    sd = VO.randomize_affine(VO.init_state_dict(cfg, seed=0), seed=1, scale=0.02)
    ref = vits.vit_base(patch_size=8, num_classes=0)
    assert sum(p.numel() for p in ref.parameters()) == VO.param_count(sd)
    ref.load_state_dict(sd, strict=True)
    ref.eval()
    x = VO.synthetic_tile(224, seed=4242, batch=1)
    with torch.no_grad():
        feat, attns, qkvs = ref.get_intermediate_feat(x, n=1)
    rows_ref = attns[0][:, :, 0, :].contiguous()                      # [1, 12, 785]
    rows_orc = VO.cls_attention_rows(sd, cfg, x)
    err = ((rows_ref - rows_orc).abs() / rows_ref.abs()).max().item()
    assert err < 1e-4, err
    att_resp, nh = ns_utils["compute_attention"](attns, 0, 28, 28, 8)
    assert nh == 12
    avg = np.mean(att_resp, axis=0)
    avg = cv2.resize(avg, (avg.shape[1] // 8, avg.shape[0] // 8))
    avg = cv2.resize(avg, (224, 224), interpolation=cv2.INTER_LINEAR)
    pil = T.ToPILImage()(x.squeeze(0)).convert("L")
    th, th2, th3 = ns_utils["threshold"](pil, avg, save=False)
    o_th, o_th2, o_th3, _, _ = PO.eval_tile(rows_ref[0].numpy(), x[0, 0].numpy(), 8)
    for a, b, what in [(o_th, th, "th"), (o_th2, th2, "th2"), (o_th3, th3, "th3")]:
        assert float((a == b).mean()) >= 0.9999, what
    g = {"x_seed": np.array(4242), "cls_rows": rows_ref.numpy(), "att_map": avg.astype(np.float32), "th": th, "th2": th2, "th3": th3,
         "feat_cls": feat[0][:, 0].numpy()}
    for k, v in checksum(sd).items():
        g["wsum/" + k] = v
    np.savez_compressed(os.path.join(OUT, "vitb8_tile.npz"), **g)
    print(f"vitb8_tile.npz: ViT-B/8 CLS rows, oracle vs reference {err:.2e}; masks agree")


if __name__ == "__main__":
    main()
