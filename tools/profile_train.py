"""Two MIM training steps at a given per-GPU batch (for ncu launch lists / captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from functools import partial
from types import SimpleNamespace as NS
import numpy as np
import torch
import vitocm_b200 as vob
from vitocm_b200 import synthetic as SY

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
torch.manual_seed(0)
enc = vob.VisionTransformerForSimMIM(patch_size=8, embed_dim=384, depth=12, num_heads=6, mlp_ratio=4, img_size=[224], qkv_bias=True,
                                     norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), precision="bf16")
mim = vob.MIM(encoder=enc, encoder_stride=8).cuda().train()
cfg = NS(TRAIN=NS(BASE_LR=5e-4, WEIGHT_DECAY=0.05, CLIP_GRAD=5.0, OPTIMIZER=NS(NAME="adamw", EPS=1e-8, BETAS=(0.9, 0.999))))
opt = vob.optimizer.build_pretrain_optimizer(cfg, mim, None)
x = SY.synthetic_tile(224, seed=1, batch=min(B, 8))
x = x[torch.arange(B) % x.shape[0]].contiguous().cuda()
rs = np.random.RandomState(0)
m = SY.random_masks(rs, B, 224, 16, 8, 0.5).cuda()
for _ in range(steps):
    opt.zero_grad()
    loss, _, _ = mim(x, m)
    loss.sum().backward()
    vob.optimizer.clip_grad_norm_(mim, 5.0)
    opt.step()
torch.cuda.synchronize()
print("loss", loss.item())
