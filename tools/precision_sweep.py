"""GPU: mask agreement and CLS-row error of every precision schedule against the fp32-parity mode (which is pinned to the
reference at 4e-6 by tests/test_gpu_parity.py), on the device's own post-processing.

    python tools/precision_sweep.py [--tiles 64] [--mosaic 4096] [--schedules bf16,fp16,fp16+mlp2,fp32] [--out file.jsonl]

Per schedule: CLS-row max / rms relative error over `tiles` synthetic 224^2 tiles, per-tile agreement of the "ours" (th) and
"heatmap" (th3) masks (mean, min, fraction of tiles >= 99.9 %), the same for the config-2 mosaic (global min-max + global Otsu),
and the segmentation time per mosaic.  No oracle import: the fp32-parity mode is the checker here."""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vitocm_b200 as vob                     # noqa: E402
from vitocm_b200 import synthetic as SY       # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tiles", type=int, default=64)
    ap.add_argument("--mosaic", type=int, default=4096)
    ap.add_argument("--arch", default="vit_small")
    ap.add_argument("--schedules", default="fp32,bf16,fp16,fp16+mlp2")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    torch.cuda.set_device(0)
    names = args.schedules.split(",")
    if names[0] != "fp32":
        names.insert(0, "fp32")
    xs = torch.cat([SY.synthetic_tile(224, seed=1234 if i == 0 else 100 + i, batch=1) for i in range(args.tiles)]).cuda()
    mosaic = torch.from_numpy(SY.synthetic_mosaic_u8(args.mosaic, seed=4321)).cuda() if args.mosaic else None
    ref = {}
    lines = []
    for name in names:
        torch.manual_seed(0)
        model = getattr(vob, args.arch)(patch_size=8, num_classes=0, precision=name, chunk_tiles=175).cuda().eval()
        out = vob.attention_masks(model, xs)
        rows, masks = out["cls_rows"].double(), out["masks"]
        rec = {"schedule": name, "arch": args.arch, "tiles": args.tiles}
        if mosaic is not None:
            seg = vob.MosaicSegmenter(model, window=224, stride=112, tile_batch=175)
            res = seg.segment(mosaic, want=("th", "th3"))
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(3):
                res = seg.segment(mosaic, want=("th", "th3"))
            torch.cuda.synchronize()
            rec["mosaic_ms"] = (time.perf_counter() - t0) / 3 * 1e3
            rec["mosaic_mp_per_s"] = res["extent"] ** 2 / 1e6 / (rec["mosaic_ms"] / 1e3)
        if name == "fp32":
            ref = {"rows": rows, "masks": masks, "res": res if mosaic is not None else None}
        rel = (rows - ref["rows"]).abs() / ref["rows"]
        rec["cls_row_max_rel_err"] = float(rel.max())
        rec["cls_row_rms_rel_err"] = float((rel ** 2).mean().sqrt())
        for i, k in ((0, "th"), (2, "th3")):
            ag = (masks[:, i] == ref["masks"][:, i]).float().mean(dim=(1, 2))
            rec[f"tile_{k}_agree_mean"], rec[f"tile_{k}_agree_min"] = float(ag.mean()), float(ag.min())
            rec[f"tile_{k}_frac_ge_999"] = float((ag >= 0.999).float().mean())
            rec[f"golden_tile_{k}"] = float(ag[0])
        if mosaic is not None:
            for k in ("th", "th3"):
                rec[f"mosaic_{k}_agree"] = float((res[k] == ref["res"][k]).float().mean())
            rec["mosaic_thresholds"] = res["thresholds"].tolist()
        lines.append(rec)
        print(json.dumps(rec), flush=True)
        del model
    if args.out:
        with open(args.out, "w") as f:
            for r in lines:
                f.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
