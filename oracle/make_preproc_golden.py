"""TEST INFRASTRUCTURE ONLY. Generates tests/golden/preproc_cases.npz by running the UNMODIFIED reference's test transform
(`get_transform_vit_default(is_train=False)`, src/pipelines/uni_pipeline.py:1233-1256 -- torchvision + Pillow underneath)
on seeded synthetic BGR images.

    python -m oracle.make_preproc_golden          # in the build container (needs /root/reference)

Small cases keep the transformed tensor itself (as the 8-bit pixels it was made from: the fp32 tensor is an exact function of
them, re-derived in the test); the full-size cases keep a SHA-256 of the fp32 tensor bytes.
"""
import hashlib
import json
import os
from types import SimpleNamespace

import numpy as np

from oracle import ref_loader

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "preproc_cases.npz")

# (crop, H, W, kind)
SMALL = [(64, 80, 100, "noise"), (64, 100, 80, "noise"), (64, 64, 64, "noise"), (64, 64, 200, "noise"), (64, 333, 129, "ramp"),
         (64, 65, 64, "noise"), (64, 128, 128, "ramp"), (64, 77, 301, "noise"), (64, 40, 50, "noise"), (64, 640, 480, "noise"),
         (32, 33, 35, "noise"), (32, 1000, 37, "ramp"), (48, 49, 51, "extreme"), (64, 97, 67, "extreme")]
FULL = [(384, 480, 640, "noise"), (384, 640, 427, "ramp"), (384, 384, 384, "noise"), (384, 1080, 1920, "noise"), (224, 375, 500, "noise")]


def make_image(h, w, kind, seed):
    rng = np.random.default_rng(seed)
    if kind == "noise":
        return rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    if kind == "extreme":                            # only 0 / 255: exercises the saturation of the cubic overshoot
        return (rng.integers(0, 2, (h, w, 3)) * 255).astype(np.uint8)
    yy, xx = np.mgrid[0:h, 0:w]
    base = np.stack([(xx * 255) // max(w - 1, 1), (yy * 255) // max(h - 1, 1), ((xx + yy) * 255) // max(h + w - 2, 1)], axis=-1)
    return np.clip(base + rng.integers(-8, 9, (h, w, 3)), 0, 255).astype(np.uint8)


def main():
    ref_loader.install_shims()
    import PIL
    import torchvision
    from src.pipelines.uni_pipeline import get_transform_vit_default

    def transform(crop):
        cfg = SimpleNamespace(test_respect_ratio_max=None, test_crop_size=crop, crop_pct=1.0)
        return get_transform_vit_default(SimpleNamespace(cfg=cfg), is_train=False)

    arrays = {}
    meta = {"pillow": PIL.__version__, "torchvision": torchvision.__version__, "small": [], "full": []}
    for i, (crop, h, w, kind) in enumerate(SMALL):
        img = make_image(h, w, kind, 1000 + i)
        ref = transform(crop)(img).numpy()                                   # float32 (3, crop, crop), RGB
        u8 = np.rint((ref * 0.5 + 0.5) * 255.0).astype(np.uint8)              # the pixels ToTensor saw (exactly recoverable)
        x = u8.astype(np.float32) / np.float32(255.0)
        assert np.array_equal(((x - np.float32(0.5)) / np.float32(0.5)), ref)
        arrays["small%d" % i] = u8
        meta["small"].append({"crop": crop, "h": h, "w": w, "kind": kind, "seed": 1000 + i})
    for i, (crop, h, w, kind) in enumerate(FULL):
        img = make_image(h, w, kind, 2000 + i)
        ref = transform(crop)(img).numpy()
        meta["full"].append({"crop": crop, "h": h, "w": w, "kind": kind, "seed": 2000 + i,
                             "sha256": hashlib.sha256(np.ascontiguousarray(ref).tobytes()).hexdigest()})
    arrays["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(OUT, **arrays)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
