"""BASELINE config 5: ViT-S/8 and ViT-B/8 CLS-attention forward at 224^2 / 448^2 / 896^2 tiles (N = 785 / 3137 / 12545):
tiles/s and algorithmic TFLOP/s of the whole forward, plus the attention kernel alone (roofline)."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vitocm_b200 as vob

def flops_per_tile(D, depth, N):
    n = N - 1
    pe = 2 * n * 192 * D
    blk = 2 * N * D * 3 * D + 4 * N * N * D + 2 * N * D * D + 16 * N * D * D
    last = 2 * N * D * D + 2 * D * D + 2 * N * D
    return pe + (depth - 1) * blk + last

rows = []
for arch, D in (("vit_small", 384), ("vit_base", 768)):
    torch.manual_seed(0)
    for size, batch in ((224, 128), (448, 32), (896, 4)):
        m = getattr(vob, arch)(patch_size=8, num_classes=0, precision="bf16", chunk_tiles=batch).cuda().eval()
        x = torch.rand(batch, 1, size, size, device="cuda").expand(-1, 3, -1, -1).contiguous()
        N = (size // 8) ** 2 + 1
        for _ in range(2):
            m.cls_attention_rows(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        e0.record()
        for _ in range(reps):
            r = m.cls_attention_rows(x)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        vob._lib.profile_enable(True)
        m.cls_attention_rows(x); torch.cuda.synchronize()
        prof = vob._lib.profile_read(); vob._lib.profile_enable(False)
        att_ms = prof["attention"][0]
        att_fl = 4 * batch * N * N * D * 11
        fl = flops_per_tile(D, 12, N) * batch
        rows.append(dict(arch=arch, tile=size, N=N, batch=batch, ms=ms, tiles_per_s=batch / ms * 1e3, tflops=fl / ms / 1e9,
                         attn_ms=att_ms, attn_tflops=att_fl / att_ms / 1e9, attn_share=att_ms / sum(v[0] for v in prof.values()),
                         rows_sum_err=float((r.sum(-1) - 1).abs().max())))
        print(json.dumps(rows[-1]), flush=True)
        del m, x
        torch.cuda.empty_cache()
