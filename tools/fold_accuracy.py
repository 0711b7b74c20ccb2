"""Accuracy of the block tail with the LayerNorm affine parameters folded into W1 / Wqkv (VITOCM_TAIL_FOLD=1) against the plain form (=0):
CLS-row error and mask agreement of the 16-bit modes against the fp32-parity mode over N synthetic 224^2 tiles (ViT-S/8).
    python tools/fold_accuracy.py [n_tiles]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import vitocm_b200 as vob  # noqa: E402
from gpu_util import build_model  # noqa: E402
from oracle import vit_oracle as VO  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
cfg = VO.ViTConfig(**VO.VIT_SMALL)
for scale in (0.02, 0.0):
    sd = VO.init_state_dict(cfg, seed=0)
    if scale > 0:
        sd = VO.randomize_affine(sd, seed=1, scale=scale)
    x = torch.cat([VO.synthetic_tile(224, seed=100 + i, batch=1) for i in range(n)]).cuda()
    m32 = build_model(cfg, sd, "fp32", chunk_tiles=64)
    ref_rows = m32.cls_attention_rows(x).double()
    ref_masks = vob.attention_masks(m32, x)["masks"]
    for precision in ("fp16", "bf16"):
        for fold in ("0", "1"):
            os.environ["VITOCM_TAIL_FOLD"] = fold
            m = build_model(cfg, sd, precision, chunk_tiles=64)
            rows = m.cls_attention_rows(x).double()
            rel = ((rows - ref_rows) / ref_rows)
            masks = vob.attention_masks(m, x)["masks"]
            out = [f"affine scale {scale} {precision} fold={fold}: rows rms {rel.pow(2).mean().sqrt().item():.3e} max {rel.abs().max().item():.3e}"]
            for i, name in ((0, "ours"), (2, "heat")):
                v = (masks[:, i] == ref_masks[:, i]).float().mean(dim=(1, 2))
                out.append(f"{name}: median {v.median().item():.5f} mean {v.mean().item():.5f} min {v.min().item():.5f} share>=.999 {(v >= 0.999).float().mean().item():.3f}")
            print(" | ".join(out), flush=True)
