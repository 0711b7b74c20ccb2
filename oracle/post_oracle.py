"""Oracle (test infrastructure, see oracle/__init__.py): numpy restatement of the reference's
attention post-processing, thresholding and sliding-window stitching.

Reference: /root/reference/Self-supervised_segmentation (abbrev. SSS)
  SSS/utils.py:55-115, 229-235, 304-317     (eval / PGT flavour)
  SSS/sw_processing.py:29-81, 113-163, 223-262 (mosaic flavour)
  SSS/eval.py:136-173                         (per-image driver loop)
Third-party arithmetic restated here because it is not under /root/reference:
  OpenCV (pinned opencv-python==4.6.0.66 in SSS/wandb/.../requirements.txt; 4.13.0 in the
  build container): cv2.resize(INTER_LINEAR) on float32 and cv2.threshold(THRESH_OTSU) on
  uint8.  oracle/make_golden.py and tests/test_oracle.py check both restatements against the
  installed cv2.
"""
from __future__ import annotations

import numpy as np

FLT_EPSILON = 1.1920928955078125e-07


# --------------------------------------------------------------------------------------
# P1: compute_attention  (SSS/utils.py:229-235)
# --------------------------------------------------------------------------------------
def compute_attention(attn0: np.ndarray, query: int, w_featmap: int, h_featmap: int, patch_size: int):
    """attn0: [B,H,N,N] (the first element of the `attentions` list).  Uses batch element 0 only.
    Returns ([H, w*p, h*p] float32 nearest-upsampled, nh)."""
    nh = attn0.shape[1]
    a = attn0[0, :, query, 1:].reshape(nh, w_featmap, h_featmap)
    a = np.repeat(np.repeat(a, patch_size, axis=1), patch_size, axis=2)   # F.interpolate(mode="nearest")
    return np.ascontiguousarray(a, dtype=np.float32), nh


def compute_attention_from_rows(cls_rows: np.ndarray, w_featmap: int, h_featmap: int, patch_size: int):
    """Same, starting from the CLS query row attn[0,:,0,:] -> [H,N]."""
    nh = cls_rows.shape[0]
    a = cls_rows[:, 1:].reshape(nh, w_featmap, h_featmap)
    a = np.repeat(np.repeat(a, patch_size, axis=1), patch_size, axis=2)
    return np.ascontiguousarray(a, dtype=np.float32), nh


def cummass_threshold(cls_rows: np.ndarray, threshold: float, w_featmap: int, h_featmap: int, patch_size: int):
    """The `--threshold` mode named by the flag at SSS/eval.py:33-34.  Its arithmetic is NOT under /root/reference (the reference
    only keeps the flag): this restates upstream DINO's visualize_attention.py (facebookresearch/dino, the file the reference's
    eval script was derived from) -- PARITY UNPINNED BY THE REFERENCE.  cls_rows [H, N] of one image ->
    (th_attn [H, w*p, h*p] float32 in {0, 1}, low-res keep mask [H, n] bool)."""
    import torch
    import torch.nn.functional as F
    a = torch.from_numpy(np.ascontiguousarray(cls_rows[:, 1:], dtype=np.float32))
    nh = a.shape[0]
    val, idx = torch.sort(a)                       # ascending
    val = val / torch.sum(val, dim=1, keepdim=True)
    cumval = torch.cumsum(val, dim=1)
    th_attn = cumval > (1 - threshold)
    idx2 = torch.argsort(idx)
    for head in range(nh):
        th_attn[head] = th_attn[head][idx2[head]]
    low = th_attn.clone().numpy()
    th_attn = th_attn.reshape(nh, w_featmap, h_featmap).float()
    up = F.interpolate(th_attn.unsqueeze(0), scale_factor=patch_size, mode="nearest")[0].numpy()
    return up, low, (cumval.numpy(), idx.numpy())


# --------------------------------------------------------------------------------------
# P3: cv2.resize(..., INTER_LINEAR) on a 2-D float32 image
# --------------------------------------------------------------------------------------
def _linear_coeffs(dst: int, src: int):
    scale = float(src) / float(dst)
    idx = np.zeros(dst, dtype=np.int64)
    frac = np.zeros(dst, dtype=np.float32)
    for d in range(dst):
        f = (d + 0.5) * scale - 0.5
        s = int(np.floor(f))
        f -= s
        if s < 0:
            s, f = 0, 0.0
        if s >= src - 1:
            s, f = src - 1, 0.0
        idx[d] = s
        frac[d] = np.float32(f)
    return idx, frac


def resize_linear(img: np.ndarray, dsize: tuple[int, int]) -> np.ndarray:
    """cv2.resize(img, dsize=(width, height), interpolation=INTER_LINEAR) for float32 [H,W]:
    half-pixel-centre bilinear, edge clamp, horizontal pass then vertical pass in float32."""
    img = np.asarray(img, dtype=np.float32)
    dw, dh = int(dsize[0]), int(dsize[1])
    sh, sw = img.shape
    xi, xf = _linear_coeffs(dw, sw)
    yi, yf = _linear_coeffs(dh, sh)
    xi1 = np.minimum(xi + 1, sw - 1)
    yi1 = np.minimum(yi + 1, sh - 1)
    a0 = (np.float32(1.0) - xf)[None, :]
    a1 = xf[None, :]
    rows = img[:, xi] * a0 + img[:, xi1] * a1                    # horizontal pass, float32
    b0 = (np.float32(1.0) - yf)[:, None]
    b1 = yf[:, None]
    out = rows[yi, :] * b0 + rows[yi1, :] * b1                   # vertical pass, float32
    return out.astype(np.float32)


# --------------------------------------------------------------------------------------
# P6: cv2.threshold(src_u8, 0, 255, THRESH_BINARY + THRESH_OTSU)
# --------------------------------------------------------------------------------------
def otsu_from_hist(hist: np.ndarray) -> int:
    """OpenCV getThreshVal_Otsu_8u scan over a 256-bin histogram, float64, first maximum wins."""
    h = np.asarray(hist, dtype=np.float64)
    total = h.sum()
    if total <= 0:
        return 0
    scale = 1.0 / total
    mu = float((np.arange(256, dtype=np.float64) * h).sum()) * scale
    mu1 = 0.0
    q1 = 0.0
    max_sigma = 0.0
    max_val = 0
    for i in range(256):
        p_i = h[i] * scale
        mu1 *= q1
        q1 += p_i
        q2 = 1.0 - q1
        if min(q1, q2) < FLT_EPSILON or max(q1, q2) > 1.0 - FLT_EPSILON:
            continue
        mu1 = (mu1 + i * p_i) / q1
        mu2 = (mu - q1 * mu1) / q2
        sigma = q1 * q2 * (mu1 - mu2) * (mu1 - mu2)
        if sigma > max_sigma:
            max_sigma = sigma
            max_val = i
    return int(max_val)


def otsu_threshold(img_u8: np.ndarray):
    """-> (t, mask) with mask = 255 where img > t else 0 (THRESH_BINARY)."""
    img_u8 = np.asarray(img_u8, dtype=np.uint8)
    hist = np.bincount(img_u8.ravel(), minlength=256)
    t = otsu_from_hist(hist)
    return t, np.where(img_u8 > t, 255, 0).astype(np.uint8)


# --------------------------------------------------------------------------------------
# P4: utils.threshold  (SSS/utils.py:55-115) -- eval / PGT flavour
# --------------------------------------------------------------------------------------
def min_max_normalize(image: np.ndarray) -> np.ndarray:
    """SSS/utils.py:55-60 and SSS/sw_processing.py:29-34: identity when the image is flat."""
    mn, mx = np.min(image), np.max(image)
    if mx == mn:
        return image
    return (image - mn) / (mx - mn)


def threshold_utils(img_u8: np.ndarray, attention: np.ndarray):
    """SSS/utils.py:62-115 with save=False.  img_u8 = np.array(PIL 'L' image).
    Returns (th 'ours', th2 'otsu on image', th3 'heatmap_threshold', result_u8, att_u8)."""
    att = min_max_normalize(np.asarray(attention))                      # :69
    img = np.array(img_u8)                                              # :71
    alpha = 0.4                                                         # :77
    att = att * 255                                                     # :78
    att_u8 = att.astype(np.uint8)                                       # :79 (truncation)
    result = (img / 2) * (1 - alpha) + (att_u8 / 2) * alpha             # :80 (float64)
    result = result.astype(np.uint8)                                    # :81
    _, th = otsu_threshold(result)                                      # :87
    t_img, _ = otsu_threshold(img)                                      # :91
    th2 = (img > float(t_img)).astype(np.uint8) * 255                   # :92-93
    _, th3 = otsu_threshold(att_u8)                                     # :95
    return th, th2, th3, result, att_u8


# --------------------------------------------------------------------------------------
# P5: sw_processing.threshold  (SSS/sw_processing.py:37-81) -- mosaic flavour
# --------------------------------------------------------------------------------------
def threshold_sw(img_u8: np.ndarray, attention: np.ndarray):
    """img_u8: gray uint8 [E,E] (the reference passes a PIL 'L' image; `img * attention` turns
    it into a uint8 ndarray times float32 -> float32).  th2 uses skimage's Otsu in the reference
    (:57, off the hot path); cv2's rule stands in for it here.
    Returns (th, th2, th3, result_u8, att_u8)."""
    att = min_max_normalize(np.asarray(attention))                      # :43
    img = np.asarray(img_u8)
    result = img * att / np.max(att)                                    # :44
    result = result.astype(np.uint8)                                    # :46
    att = att * 255                                                     # :47
    att_u8 = att.astype(np.uint8)                                       # :48
    _, th = otsu_threshold(result)                                      # :54
    t_img, _ = otsu_threshold(img)                                      # :57-59 (see docstring)
    th2 = (img > t_img).astype(np.uint8) * 255
    _, th3 = otsu_threshold(att_u8)                                     # :61
    return th, th2, th3, result, att_u8


# --------------------------------------------------------------------------------------
# T1: sliding_window (SSS/sw_processing.py:151-163) on a numpy image [H,W] or [H,W,C]
# --------------------------------------------------------------------------------------
def sliding_window_origins(size: int, stride: int) -> list[int]:
    """range(0, size - 2*stride, stride) -- the reference's loop bound (:156-157)."""
    return list(range(0, size - stride * 2, stride))


def sliding_window(image: np.ndarray, stride: int = 128, window_size: int = 384) -> list[np.ndarray]:
    """PIL's crop zero-pads beyond the image; so does this."""
    image = np.asarray(image)
    hgt, wid = image.shape[0], image.shape[1]
    # reference: `height, width = image.size` is PIL's (W, H); y loops over `height`=W, x over `width`=H
    crops = []
    for y in sliding_window_origins(wid, stride):
        for x in sliding_window_origins(hgt, stride):
            win = np.zeros((window_size, window_size) + image.shape[2:], dtype=image.dtype)
            ys, xs = min(y + window_size, hgt), min(x + window_size, wid)
            if ys > y and xs > x:
                win[: ys - y, : xs - x] = image[y:ys, x:xs]
            crops.append(win)
    return crops


# --------------------------------------------------------------------------------------
# T3: concat_crops with linear-ramp blending (SSS/sw_processing.py:113-149)
# --------------------------------------------------------------------------------------
def _blend_h(left: np.ndarray, right: np.ndarray) -> np.ndarray:
    """:143-149 -- float64 weights, result stored into zeros_like(left) (dtype of the crops)."""
    w = np.linspace(1, 0, left.shape[1])
    shape = (1, -1) + (1,) * (left.ndim - 2)
    w = w.reshape(shape)
    return (left * w + right * (1 - w)).astype(left.dtype)


def _blend_v(top: np.ndarray, bottom: np.ndarray) -> np.ndarray:
    """:136-141."""
    w = np.linspace(1, 0, top.shape[0])
    shape = (-1,) + (1,) * (top.ndim - 1)
    w = w.reshape(shape)
    return (top * w + bottom * (1 - w)).astype(top.dtype)


def concat_crops_blend(crops: list[np.ndarray], stride: int, window_size: int) -> np.ndarray:
    """:113-134, vectorised per seam (same arithmetic, same accumulation dtype and order)."""
    it = int(np.sqrt(len(crops)))
    step = window_size - stride
    vertical = None
    for i in range(it):
        horizontal = crops[i * it]
        for j in range(1, it):
            right = crops[i * it + j]
            overlap = _blend_h(horizontal[:, -step:], right[:, :-stride])
            horizontal = np.concatenate((horizontal[:, :-step], overlap, right[:, -stride:]), axis=1)
        if i == 0:
            vertical = horizontal
        else:
            top = _blend_v(vertical[-step:, :], horizontal[:-stride, :])
            vertical = np.concatenate((vertical[:-step, :], top, horizontal[-stride:, :]), axis=0)
    return vertical


def concat_crops_plain(crops: list[np.ndarray]) -> np.ndarray:
    """SSS/utils.py:304-317 -- non-overlapping row-major tiling (eval.py --crop 4|16)."""
    it = int(np.sqrt(len(crops)))
    rows = [np.concatenate(crops[i * it:(i + 1) * it], axis=1) for i in range(it)]
    return np.concatenate(rows, axis=0)


def blend_profiles(n: int, stride: int, window_size: int) -> np.ndarray:
    """Separable weights of concat_crops_blend: out(y,x) = sum_ij P[i,y] P[j,x] tile_ij(y-iS, x-jS)
    (SURVEY.md section 7).  Returns P [n, E] float64, derived by pushing one-hot rows through the
    sequential blend so that it holds for every (W, S), including the asymmetric W=3S case."""
    E = (n - 1) * stride + window_size
    P = np.zeros((n, E), dtype=np.float64)
    step = window_size - stride
    for k in range(n):
        cur = np.full(window_size, 1.0 if k == 0 else 0.0)
        for j in range(1, n):
            right = np.full(window_size, 1.0 if k == j else 0.0)
            w = np.linspace(1, 0, step)
            ov = cur[-step:] * w + right[:-stride] * (1 - w)
            cur = np.concatenate((cur[:-step], ov, right[-stride:]))
        P[k] = cur
    return P


def concat_crops_overlap(crops: list[np.ndarray], stride: int) -> np.ndarray:
    """utils.py:319-347.  n x n crops of size W that overlap by V = 2*stride: the running image and the
    next crop are floor-halved and added in the overlap (`a // 2 + b // 2`, in the crops' dtype), first
    along x (:324-332), then along y (:334-345) -- where the LAST strip (:337-339) is appended without
    blending, so its overlap rows keep the running image."""
    n = int(np.sqrt(len(crops)))
    V = 2 * stride
    W = crops[0].shape[0]
    step = W - V

    def halves(a, b):
        return a // 2 + b // 2

    strips = []
    for i in range(n):
        strip = np.array(crops[i * n])
        for j in range(1, n):
            nxt = np.asarray(crops[i * n + j])
            x0 = j * step                                  # the running strip is x0 + V wide
            grown = np.empty((strip.shape[0], x0 + W) + strip.shape[2:], dtype=strip.dtype)
            grown[:, :x0] = strip[:, :x0]
            grown[:, x0:x0 + V] = halves(strip[:, x0:x0 + V], nxt[:, :V])
            grown[:, x0 + V:] = nxt[:, V:]
            strip = grown
        strips.append(strip)
    out = strips[0]
    for i in range(1, n):
        y0 = i * step
        grown = np.empty((y0 + W,) + out.shape[1:], dtype=out.dtype)
        grown[:y0] = out[:y0]
        grown[y0:y0 + V] = out[y0:y0 + V] if i == n - 1 else halves(out[y0:y0 + V], strips[i][:V])
        grown[y0 + V:] = strips[i][V:]
        out = grown
    return out


def sliding_window_utils(image: np.ndarray, window_size: int, stride: int) -> list[np.ndarray]:
    """utils.py:349-362: the same window walk as sw_processing.sliding_window with (window_size, stride) swapped."""
    return sliding_window(image, stride, window_size)


def eval_cropped(cls_rows_crops: np.ndarray, crops01: np.ndarray, patch_size: int = 8):
    """eval.py:145-173 (`--crop 4|16`).  cls_rows_crops [cr*cr, H, N]: CLS rows of every crop of ONE image, row-major;
    crops01 [cr*cr, s, s]: channel 0 of the crops (float in [0,1]).  Per crop: head mean of the nearest-upsampled
    CLS maps (median_filter size 1 = identity); plain concat (:160-161); resize /p then up to the image size
    (:169-171); threshold against the PIL "L" image of the tiled crops (:173).
    Returns (attention [S,S] float32, threshold_utils tuple)."""
    ncrop, s = crops01.shape[0], crops01.shape[-1]
    f = s // patch_size
    maps = []
    for c in range(ncrop):
        a, _ = compute_attention_from_rows(cls_rows_crops[c], f, f, patch_size)
        maps.append(np.mean(a, axis=0))                                      # :156
    avg = concat_crops_plain(maps)                                           # :160
    img = concat_crops_plain([np.asarray(c, dtype=np.float32) for c in crops01])   # :161
    small = resize_linear(avg, (avg.shape[1] // patch_size, avg.shape[0] // patch_size))   # :169
    att = resize_linear(small, (img.shape[-1], img.shape[-1]))               # :171
    img_u8 = np.floor(img * np.float32(255.0)).astype(np.uint8)              # ToPILImage + convert("L") on R=G=B
    return att, threshold_utils(img_u8, att)


# --------------------------------------------------------------------------------------
# drivers
# --------------------------------------------------------------------------------------
def tile_attention_map(cls_rows: np.ndarray, size: int, patch_size: int) -> np.ndarray:
    """eval.py:140-171 for one tile: head mean of the nearest-upsampled CLS maps, median_filter
    size 1 (identity), cv2.resize down by p then up INTER_LINEAR -> [S,S] float32."""
    f = size // patch_size
    a, _ = compute_attention_from_rows(cls_rows, f, f, patch_size)
    avg = np.mean(a, axis=0)                                            # eval.py:142
    small = resize_linear(avg, (avg.shape[1] // patch_size, avg.shape[0] // patch_size))   # :169
    return resize_linear(small, (size, size))                           # :171


def eval_tile(cls_rows: np.ndarray, gray01: np.ndarray, patch_size: int = 8):
    """eval.py:136-173 (`--crop 1`, method ours/otsu/heatmap_threshold) for one image.
    gray01: [S,S] float in [0,1] (channel 0 of the R=G=B input tensor); PIL 'L' of ToPILImage
    is floor(x*255) for gray input.  Returns the threshold_utils tuple."""
    size = gray01.shape[0]
    att = tile_attention_map(cls_rows, size, patch_size)
    img_u8 = np.floor(np.asarray(gray01, dtype=np.float32) * np.float32(255.0)).astype(np.uint8)
    return threshold_utils(img_u8, att)


def sw_tile_map(cls_rows: np.ndarray, window: int, patch_size: int = 8) -> np.ndarray:
    """sw_processing.py:243-257 for one crop: head mean -> per-tile min-max*255 (float32) ->
    resize //8 -> resize x8 INTER_LINEAR."""
    f = window // patch_size
    a, _ = compute_attention_from_rows(cls_rows, f, f, patch_size)
    avg = np.mean(a, axis=0)
    avg = (avg - avg.min()) / (avg.max() - avg.min())                  # :253
    avg = avg * 255                                                     # :254
    small = resize_linear(avg, (avg.shape[1] // 8, avg.shape[0] // 8))  # :255
    return resize_linear(small, (small.shape[0] * 8, small.shape[0] * 8))   # :257


def mosaic_segment(cls_rows_all: np.ndarray, mosaic_u8: np.ndarray, stride: int, window: int, patch_size: int = 8):
    """sw_processing.py:223-262.  cls_rows_all: [T,H,N] for the T = n*n crops in row-major
    order; mosaic_u8: gray [E0,E0].  Returns (stitched float32 [E,E], threshold_sw tuple)."""
    n = int(np.sqrt(cls_rows_all.shape[0]))
    maps = [sw_tile_map(cls_rows_all[t], window, patch_size) for t in range(n * n)]
    stitched = concat_crops_blend(maps, stride, window)                 # :259
    crops = sliding_window(mosaic_u8, stride, window)
    gray = concat_crops_blend(crops, stride, window)                    # :225 (uint8 accumulate)
    return stitched, threshold_sw(gray, stitched), gray
