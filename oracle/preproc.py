"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

CPU restatement (numpy, integer arithmetic + IEEE doubles) of the reference's TEST-time image transform
(`get_transform_vit_default(is_train=False)`, src/pipelines/uni_pipeline.py:1233-1256):

    BGR2RGB (src/data_layer/transform.py:47-50) -> ToPILImage -> Resize(floor(crop / crop_pct), BICUBIC) -> CenterCrop(crop)
    -> ToTensor -> Normalize(0.5, 0.5)

The resampling arithmetic itself is NOT in /root/reference: it lives in two third-party dependencies the reference imports,
  * torchvision (README.md:20 pins 0.7.0; 0.26.0 runs here): `Resize` with an int size scales the SHORTER edge to `size`
    and the longer one to int(size * long / short); `CenterCrop` takes top/left = int(round((full - crop) / 2.0))
    (Python round: half to even);
  * Pillow (unpinned by the reference; 12.2.0 runs here): `Image.resize(..., BICUBIC)` = ImagingResample for 8-bit
    channels: a horizontal pass followed by a vertical pass (each skipped when that dimension is unchanged), 8-bit
    intermediate, per-output-pixel coefficient windows computed in double precision (Keys cubic, a = -0.5, support
    2 * max(scale, 1)), normalised, converted to 22-bit fixed point, accumulated in int32 from 1 << 21 and shifted back
    with saturation to 0..255.
Their published algorithm is restated below; parity is pinned against torchvision + Pillow run through the reference's own
`get_transform_vit_default` in the build container (oracle/make_preproc_golden.py -> tests/golden/preproc_*.npz).
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2          # Pillow: coefficients are 22-bit fixed point


def bicubic_filter(x):
    """Keys cubic convolution kernel with a = -0.5 (Pillow's BICUBIC), evaluated with Pillow's operation order."""
    x = np.abs(np.asarray(x, dtype=np.float64))
    a = -0.5
    near = ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    far = (((x - 5) * x + 8) * x - 4) * a
    return np.where(x < 1.0, near, np.where(x < 2.0, far, 0.0))


def precompute_coeffs(in_size, out_size):
    """Per output coordinate: first source index, tap count and the normalised double weights (Pillow precompute_coeffs with
    the box = the whole axis). Returns (ksize, xmin int[out], count int[out], kk float64[out, ksize])."""
    scale = float(in_size) / out_size
    filterscale = scale if scale >= 1.0 else 1.0
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    xmin = np.zeros(out_size, dtype=np.int64)
    cnt = np.zeros(out_size, dtype=np.int64)
    kk = np.zeros((out_size, ksize), dtype=np.float64)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        lo = int(center - support + 0.5)            # C cast: truncation toward zero
        if lo < 0:
            lo = 0
        hi = int(center + support + 0.5)
        if hi > in_size:
            hi = in_size
        n = hi - lo
        ww = 0.0
        w = np.zeros(ksize, dtype=np.float64)
        for x in range(n):
            w[x] = float(bicubic_filter((x + lo - center + 0.5) * ss))
            ww += w[x]
        if ww != 0.0:
            for x in range(n):
                w[x] /= ww
        xmin[xx], cnt[xx], kk[xx] = lo, n, w
    return ksize, xmin, cnt, kk


def normalize_coeffs_8bpc(kk):
    """double weights -> int32 fixed point, rounding half away from zero by a C cast (Pillow normalize_coeffs_8bpc)."""
    v = kk * float(1 << PRECISION_BITS)
    return np.where(kk < 0, np.trunc(-0.5 + v), np.trunc(0.5 + v)).astype(np.int64)


def _clip8(acc):
    return np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)       # arithmetic shift = floor, then saturate


def _resample_axis(img, out_size, axis):
    """One pass of ImagingResample over `axis` (0 = rows / vertical, 1 = columns / horizontal) of an (H, W, C) uint8 array."""
    in_size = img.shape[axis]
    ksize, xmin, cnt, kk = precompute_coeffs(in_size, out_size)
    ik = normalize_coeffs_8bpc(kk)
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((out_size,) + src.shape[1:], dtype=np.uint8)
    for xx in range(out_size):
        acc = np.full(src.shape[1:], 1 << (PRECISION_BITS - 1), dtype=np.int64)
        for x in range(int(cnt[xx])):
            acc += src[xmin[xx] + x] * ik[xx, x]
        out[xx] = _clip8(acc)
    return np.moveaxis(out, 0, axis)


def pil_resize_bicubic(img, out_w, out_h):
    """Image.resize((out_w, out_h), BICUBIC) on an (H, W, C) uint8 array: horizontal pass first, then vertical, each only if
    that dimension changes (Pillow ImagingResample)."""
    if img.shape[1] != out_w:
        img = _resample_axis(img, out_w, 1)
    if img.shape[0] != out_h:
        img = _resample_axis(img, out_h, 0)
    return img


def resized_size(h, w, size):
    """torchvision Resize with an int: shorter edge -> size, longer edge -> int(size * long / short)."""
    short, long = (w, h) if w <= h else (h, w)
    new_short, new_long = size, int(size * long / short)
    return (new_long, new_short) if w <= h else (new_short, new_long)      # (new_h, new_w)


def center_crop_origin(full, crop):
    return int(round((full - crop) / 2.0))                                  # Python round: half to even


def test_transform_u8(img_bgr, crop_size, crop_pct=1.0):
    """(H, W, 3) uint8 BGR -> the 8-bit pixels after Resize + CenterCrop, (crop, crop, 3), channel order unchanged (the channel
    flip commutes with the per-channel resampling and is applied with ToTensor/Normalize downstream)."""
    h, w = img_bgr.shape[:2]
    size = int(math.floor(crop_size / crop_pct))
    nh, nw = resized_size(h, w, size)
    if nh < crop_size or nw < crop_size:
        raise ValueError("resized image smaller than the crop (torchvision would zero-pad); not on the reference's path")
    r = pil_resize_bicubic(img_bgr, nw, nh)
    top, left = center_crop_origin(nh, crop_size), center_crop_origin(nw, crop_size)
    return np.ascontiguousarray(r[top:top + crop_size, left:left + crop_size])


def test_transform(img_bgr, crop_size, crop_pct=1.0):
    """The whole reference transform: float32 (3, crop, crop) RGB, ((x / 255) - 0.5) / 0.5 in fp32 like torchvision."""
    u8 = test_transform_u8(img_bgr, crop_size, crop_pct)[:, :, ::-1]
    x = u8.astype(np.float32).transpose(2, 0, 1) / np.float32(255.0)
    return ((x - np.float32(0.5)) / np.float32(0.5)).astype(np.float32)
