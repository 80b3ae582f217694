"""The numpy restatement of the reference's test-time image transform (oracle/preproc.py) against the goldens produced by the
reference's own `get_transform_vit_default` (torchvision + Pillow) -- oracle/make_preproc_golden.py."""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import preproc
from oracle.make_preproc_golden import make_image

from .helpers import GOLDEN_DIR


def _cases():
    z = np.load(os.path.join(GOLDEN_DIR, "preproc_cases.npz"))
    return z, json.loads(bytes(z["meta"]).decode())


def test_small_cases_bit_exact():
    z, meta = _cases()
    for i, c in enumerate(meta["small"]):
        img = make_image(c["h"], c["w"], c["kind"], c["seed"])
        got = preproc.test_transform_u8(img, c["crop"])[:, :, ::-1].transpose(2, 0, 1)       # RGB, CHW like the golden
        assert np.array_equal(got, z["small%d" % i]), c


@pytest.mark.parametrize("idx", [0, 1, 2, 4])
def test_full_size_cases_by_digest(idx):
    _, meta = _cases()
    c = meta["full"][idx]
    img = make_image(c["h"], c["w"], c["kind"], c["seed"])
    ref = preproc.test_transform(img, c["crop"])
    assert hashlib.sha256(np.ascontiguousarray(ref).tobytes()).hexdigest() == c["sha256"], c


def test_geometry_matches_torchvision_rules():
    assert preproc.resized_size(480, 640, 384) == (384, 512)
    assert preproc.resized_size(640, 480, 384) == (512, 384)
    assert preproc.resized_size(333, 129, 64) == (165, 64)
    assert preproc.center_crop_origin(165, 64) == 50          # 50.5 rounds to even
    assert preproc.center_crop_origin(167, 64) == 52          # 51.5 rounds to even
    with pytest.raises(ValueError):
        preproc.test_transform_u8(np.zeros((10, 10, 3), np.uint8), 64, crop_pct=2.0)
