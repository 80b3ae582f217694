"""Device-side Resize(BICUBIC) + CenterCrop (vc_resize_crop_u8, through the C ABI) against the oracle and the reference goldens:
bit-exact on every byte."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from oracle import preproc
from oracle.make_preproc_golden import make_image

from .helpers import GOLDEN_DIR

pytestmark = pytest.mark.gpu


def _cases():
    z = np.load(os.path.join(GOLDEN_DIR, "preproc_cases.npz"))
    return z, json.loads(bytes(z["meta"]).decode())


def test_small_goldens_ragged_batches():
    from vitcap_b200.preproc import DeviceTestTransform
    z, meta = _cases()
    by_crop = {}
    for i, c in enumerate(meta["small"]):
        by_crop.setdefault(c["crop"], []).append((i, c))
    for crop, lst in by_crop.items():
        imgs = [make_image(c["h"], c["w"], c["kind"], c["seed"]) for _, c in lst]
        out = DeviceTestTransform(crop)(imgs).cpu().numpy()                   # one ragged batch per crop size
        for (i, c), o in zip(lst, out):
            assert np.array_equal(o[:, :, ::-1].transpose(2, 0, 1), z["small%d" % i]), c


def test_full_size_goldens_through_patch_embed_transform():
    """uint8 resize on the device, then the fp32 tensor ToTensor/Normalize would give: digest equals the reference's."""
    from vitcap_b200.preproc import DeviceTestTransform
    _, meta = _cases()
    for c in meta["full"]:
        img = make_image(c["h"], c["w"], c["kind"], c["seed"])
        u8 = DeviceTestTransform(c["crop"])([img]).cpu().numpy()[0]
        x = u8[:, :, ::-1].astype(np.float32).transpose(2, 0, 1) / np.float32(255.0)
        ref = (x - np.float32(0.5)) / np.float32(0.5)
        assert hashlib.sha256(np.ascontiguousarray(ref).tobytes()).hexdigest() == c["sha256"], c


def test_random_ragged_batch_vs_oracle():
    from vitcap_b200.preproc import DeviceTestTransform
    rng = np.random.default_rng(5)
    shapes = [(int(rng.integers(96, 400)), int(rng.integers(96, 400))) for _ in range(12)] + [(96, 96), (96, 1500), (2000, 97)]
    imgs = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in shapes]
    out = DeviceTestTransform(96)(imgs).cpu().numpy()
    for im, o in zip(imgs, out):
        assert np.array_equal(o, preproc.test_transform_u8(im, 96)), im.shape


def test_identity_and_crop_only():
    from vitcap_b200.preproc import DeviceTestTransform
    rng = np.random.default_rng(6)
    a = rng.integers(0, 256, (128, 128, 3), dtype=np.uint8)
    b = rng.integers(0, 256, (128, 301, 3), dtype=np.uint8)                  # already 128 high: pure center crop
    out = DeviceTestTransform(128)([a, b]).cpu().numpy()
    assert np.array_equal(out[0], a)
    left = int(round((301 - 128) / 2.0))
    assert np.array_equal(out[1], b[:, left:left + 128])


def test_feeds_the_caption_model_like_the_host_transform():
    """uint8 device transform -> model == reference-style host transform (fp32 tensor) -> model, bit for bit (exact mode)."""
    from vitcap_b200 import config as vcfg, synth
    from vitcap_b200.model import FastImageCaptioning
    from vitcap_b200.preproc import DeviceTestTransform
    cfg = vcfg.tiny()
    S = cfg.img_size
    rng = np.random.default_rng(9)
    imgs = [rng.integers(0, 256, (int(rng.integers(S, 3 * S)), int(rng.integers(S, 3 * S)), 3), dtype=np.uint8) for _ in range(3)]
    host = torch.from_numpy(np.stack([preproc.test_transform(im, S) for im in imgs]))
    m = FastImageCaptioning(cfg, mode="fp32")
    m.load_state_dict(synth.make_state_dict(cfg, seed=3))
    m = m.cuda()
    data = synth.make_text_inputs(cfg, 3)
    data = {k: v.cuda() for k, v in data.items()}
    ids_a, lp_a = m(dict(data, image=host.cuda()))
    ids_b, lp_b = m(dict(data, image=DeviceTestTransform(S)(imgs)))
    assert torch.equal(ids_a, ids_b) and torch.equal(lp_a, lp_b)


def test_plan_refuses_bad_geometry():
    from vitcap_b200 import ops
    with pytest.raises(RuntimeError):
        ops.resize_crop_plan(torch.tensor([[0, 5]], dtype=torch.int32), 64, 64)
    with pytest.raises(RuntimeError):
        ops.resize_crop_plan(torch.tensor([[50, 50]], dtype=torch.int32), 32, 64)       # resize_to < crop


def test_pipelined_calls_do_not_overwrite_a_pending_upload():
    """The transform's uploads are asynchronous: while the GPU is still busy with earlier work the host may already be packing
    the next batches. The pinned staging buffers are rotated and guarded by events, so a batch queued behind a long-running
    kernel is still resized from ITS pixels (a single unguarded buffer would be overwritten by the following call)."""
    from vitcap_b200.preproc import DeviceTestTransform
    rng = np.random.default_rng(11)
    batches = [[rng.integers(0, 256, (int(rng.integers(100, 260)), int(rng.integers(100, 260)), 3), dtype=np.uint8) for _ in range(4)]
               for _ in range(5)]
    t = DeviceTestTransform(96)
    ref = [t(b).cpu().numpy() for b in batches]                  # serial: every call is drained before the next
    torch.cuda.synchronize()
    torch.cuda._sleep(int(1.5e9))                                # ~1 s of GPU work ahead of the first upload
    outs = [t(b) for b in batches]                               # the host runs ahead of the device
    torch.cuda.synchronize()
    for r, o in zip(ref, outs):
        assert np.array_equal(r, o.cpu().numpy())
