"""ORACLE (test infrastructure, not product code) — CPU restatement of the gray256 stage of the hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.

The reference builds the 256x256 gray image every frequency feature is computed from with two third-party libraries
(train_fusion_head_only.py:142-148, deepfake-detector-v2/app.py:736-749):

    ImageOps.exif_transpose(pil).convert("L")            Pillow   (ITU-R 601-2 luma, integer)
    cv2.createCLAHE(2.0, (8, 8)).apply(u8)               OpenCV   (always in the trainers, optional in the apps)
    Image.fromarray(arr).resize((256, 256), BICUBIC)     Pillow   (antialiased separable resample, 22-bit fixed point)
    np.float32 / 255

Neither library is vendored or pinned by the reference (requirements.txt lists bare `Pillow`, `opencv-python`); the
algorithms below restate their published C sources and are PINNED against the versions installed in the build
container (Pillow 12.2.0, opencv-python-headless 4.13.0) by tests/test_oracle_cpu.py (bit-exact on random and
structured images at several sizes) and by tests/golden/gray_golden.npz, which oracle/make_golden.py wrote by calling
the libraries exactly as the reference does.

  Pillow  src/libImaging/Convert.c   rgb2l:    L = (R*19595 + G*38470 + B*7471 + 0x8000) >> 16
  Pillow  src/libImaging/Resample.c  precompute_coeffs / normalize_coeffs_8bpc / ImagingResampleHorizontal_8bpc /
                                     ImagingResampleVertical_8bpc (horizontal pass first, u8 between the passes)
  OpenCV  modules/imgproc/src/clahe.cpp  CLAHE_CalcLut_Body<uchar,256,0> / CLAHE_Interpolation_Body<uchar,0>
"""
from __future__ import annotations

import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2  # Resample.c


def luma_u8(rgb: np.ndarray) -> np.ndarray:
    """[..., 3] u8 -> [...] u8, Pillow 'RGB' -> 'L'."""
    r, g, b = (rgb[..., i].astype(np.int64) for i in range(3))
    return ((r * 19595 + g * 38470 + b * 7471 + 0x8000) >> 16).astype(np.uint8)


def _bicubic(x: float) -> float:
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def _bilinear(x: float) -> float:
    if x < 0.0:
        x = -x
    if x < 1.0:
        return 1.0 - x
    return 0.0


def resample_coeffs(in_size: int, out_size: int, filter: str = "bicubic"):
    """Pillow precompute_coeffs + normalize_coeffs_8bpc for the bicubic (support 2) / bilinear (support 1) filter over the
    full image.  Returns (xmin int32[out], count int32[out], kk int32[out, ksize])."""
    filt, supp = (_bicubic, 2.0) if filter == "bicubic" else (_bilinear, 1.0)
    scale = filterscale = in_size / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = supp * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    xmin_a = np.zeros(out_size, np.int32)
    cnt_a = np.zeros(out_size, np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        ww = 0.0
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        k = [0.0] * ksize
        for x in range(xmax):
            w = filt((x + xmin - center + 0.5) * ss)
            k[x] = w
            ww += w
        for x in range(xmax):
            if ww != 0.0:
                k[x] /= ww
        for x in range(ksize):
            v = k[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        xmin_a[xx], cnt_a[xx] = xmin, xmax
    return xmin_a, cnt_a, kk


def _resample_last_axis(img: np.ndarray, out_size: int, filter: str = "bicubic") -> np.ndarray:
    """One Pillow 8bpc pass along the last axis: u8 [..., n] -> u8 [..., out_size]."""
    n = img.shape[-1]
    xmin, cnt, kk = resample_coeffs(n, out_size, filter)
    src = img.astype(np.int64)
    out = np.empty(img.shape[:-1] + (out_size,), np.uint8)
    for xx in range(out_size):
        c = int(cnt[xx])
        acc = (src[..., xmin[xx]:xmin[xx] + c] * kk[xx, :c].astype(np.int64)).sum(-1) + (1 << (PRECISION_BITS - 1))
        out[..., xx] = np.clip(acc >> PRECISION_BITS, 0, 255)
    return out


def resize_bicubic_u8(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """Pillow Image.resize((out_w, out_h), BICUBIC) of an 'L' image [H, W] u8 (a pass is skipped when its size is
    unchanged, as ImagingResample does)."""
    h, w = img.shape
    if w != out_w:
        img = _resample_last_axis(img, out_w)
    if h != out_h:
        img = np.ascontiguousarray(_resample_last_axis(np.ascontiguousarray(img.T), out_h).T)
    return img


def resize_u8(img: np.ndarray, out_h: int, out_w: int, filter: str = "bilinear") -> np.ndarray:
    """Pillow Image.resize((out_w, out_h), BILINEAR|BICUBIC) of an 8-bit image [H, W] or [H, W, C]: every band is
    resampled independently with the same coefficients, horizontal pass first."""
    a = img if img.ndim == 3 else img[..., None]
    h, w, _ = a.shape
    if w != out_w:
        a = np.moveaxis(_resample_last_axis(np.ascontiguousarray(np.moveaxis(a, 1, -1)), out_w, filter), -1, 1)
    if h != out_h:
        a = np.moveaxis(_resample_last_axis(np.ascontiguousarray(np.moveaxis(a, 0, -1)), out_h, filter), -1, 0)
    a = np.ascontiguousarray(a)
    return a if img.ndim == 3 else a[..., 0]


def _reflect101(i: int, n: int) -> int:
    if n == 1:
        return 0
    while i < 0 or i >= n:
        i = -i if i < 0 else 2 * (n - 1) - i
    return i


def clahe_u8(img: np.ndarray, clip_limit: float = 2.0, tiles: int = 8) -> np.ndarray:
    """cv2.createCLAHE(clip_limit, (tiles, tiles)).apply(img) for an 8-bit single-channel image."""
    h, w = img.shape
    if w % tiles == 0 and h % tiles == 0:
        ext, tw, th = img, w // tiles, h // tiles
    else:
        eh, ew = h + (tiles - h % tiles), w + (tiles - w % tiles)  # clahe.cpp pads by a whole tile step when h % tiles == 0
        ys = [_reflect101(y, h) for y in range(eh)]
        xs = [_reflect101(x, w) for x in range(ew)]
        ext = img[np.ix_(ys, xs)]
        tw, th = ew // tiles, eh // tiles
    area = tw * th
    lut_scale = np.float32(255.0) / np.float32(area)
    clip = 0
    if clip_limit > 0.0:
        clip = max(int(clip_limit * area / 256), 1)
    luts = np.zeros((tiles, tiles, 256), np.uint8)
    for ty in range(tiles):
        for tx in range(tiles):
            hist = np.bincount(ext[ty * th:(ty + 1) * th, tx * tw:(tx + 1) * tw].ravel(), minlength=256).astype(np.int64)
            if clip > 0:
                clipped = int(np.maximum(hist - clip, 0).sum())
                hist = np.minimum(hist, clip)
                batch = clipped // 256
                residual = clipped - batch * 256
                hist += batch
                if residual != 0:
                    step = max(256 // residual, 1)
                    i = 0
                    while i < 256 and residual > 0:
                        hist[i] += 1
                        i += step
                        residual -= 1
            csum = np.cumsum(hist).astype(np.float32) * lut_scale  # int -> float conversion, one fp32 multiply
            luts[ty, tx] = np.clip(np.rint(csum), 0, 255).astype(np.uint8)  # saturate_cast<uchar>: cvRound (half to even)
    inv_tw, inv_th = np.float32(1.0) / np.float32(tw), np.float32(1.0) / np.float32(th)
    xf = np.arange(w, dtype=np.float32) * inv_tw - np.float32(0.5)
    tx1 = np.floor(xf).astype(np.int32)
    xa = (xf - tx1.astype(np.float32)).astype(np.float32)
    xa1 = (np.float32(1.0) - xa).astype(np.float32)
    tx2 = np.minimum(tx1 + 1, tiles - 1)
    tx1 = np.maximum(tx1, 0)
    yf = np.arange(h, dtype=np.float32) * inv_th - np.float32(0.5)
    ty1 = np.floor(yf).astype(np.int32)
    ya = (yf - ty1.astype(np.float32)).astype(np.float32)
    ya1 = (np.float32(1.0) - ya).astype(np.float32)
    ty2 = np.minimum(ty1 + 1, tiles - 1)
    ty1 = np.maximum(ty1, 0)
    v = img.astype(np.intp)
    l11 = luts[ty1[:, None], tx1[None, :], v].astype(np.float32)
    l12 = luts[ty1[:, None], tx2[None, :], v].astype(np.float32)
    l21 = luts[ty2[:, None], tx1[None, :], v].astype(np.float32)
    l22 = luts[ty2[:, None], tx2[None, :], v].astype(np.float32)
    # fp32 throughout, in clahe.cpp's order, no fused multiply-add
    top = (l11 * xa1[None, :]).astype(np.float32) + (l12 * xa[None, :]).astype(np.float32)
    bot = (l21 * xa1[None, :]).astype(np.float32) + (l22 * xa[None, :]).astype(np.float32)
    res = (top.astype(np.float32) * ya1[:, None]).astype(np.float32) + (bot.astype(np.float32) * ya[:, None]).astype(np.float32)
    return np.clip(np.rint(res.astype(np.float32)), 0, 255).astype(np.uint8)


def gray256_from_rgb_u8(rgb_u8_hwc: np.ndarray, clahe: bool) -> np.ndarray:
    """The whole stage: [H, W, 3] u8 -> [256, 256] f32 in [0, 1]."""
    g = luma_u8(rgb_u8_hwc)
    if clahe:
        g = clahe_u8(g)
    g = resize_bicubic_u8(g, 256, 256)
    return g.astype(np.float32) / np.float32(255.0)


# ---------------------------------------------------------------------------------------------------------
# seeded test images shared by oracle/make_golden.py (which stores the libraries' outputs) and tests/ (which
# regenerate the inputs)
# ---------------------------------------------------------------------------------------------------------
GOLDEN_CASES = [  # (height, width, kind, seed)
    (384, 384, "noise", 1), (224, 224, "waves", 2), (256, 256, "noise", 3), (300, 451, "waves", 4),
    (97, 64, "noise", 5), (512, 384, "edges", 6), (33, 1000, "waves", 7), (768, 1024, "edges", 8),
]


def synthetic_rgb(h: int, w: int, kind: str, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    if kind == "noise":
        return rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    yy, xx = np.mgrid[0:h, 0:w]
    if kind == "waves":  # smooth gradients: exercises CLAHE's clipping / redistribution and rounding ties
        base = (127 + 120 * np.sin(xx / (11.0 + seed)) * np.cos(yy / (19.0 + seed))).astype(np.uint8)
        return np.stack([base, np.roll(base, 5, 1), 255 - base], -1)
    img = np.zeros((h, w, 3), np.uint8)  # "edges": flat regions, hard steps and a little noise
    img[h // 3:, w // 2:] = 255
    img[: h // 4, : w // 3, 1] = 90
    img[::7, ::5] = rng.integers(0, 256, img[::7, ::5].shape, dtype=np.uint8)
    return img
