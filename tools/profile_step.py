"""Small fixed workload for ncu: ViT-S/8 CLS-attention forward on one chunk of 32 synthetic tiles, twice."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vitocm_b200 as vob  # noqa: E402

tiles = int(sys.argv[1]) if len(sys.argv) > 1 else 32
arch = sys.argv[2] if len(sys.argv) > 2 else "vit_small"
precision = sys.argv[3] if len(sys.argv) > 3 else "bf16"
torch.manual_seed(0)
m = getattr(vob, arch)(patch_size=8, num_classes=0, precision=precision, chunk_tiles=tiles).cuda().eval()
x = torch.rand(tiles, 1, 224, 224, device="cuda").expand(-1, 3, -1, -1).contiguous()
for _ in range(2):
    rows = m.cls_attention_rows(x)
torch.cuda.synchronize()
print("ok", float(rows.sum()))
