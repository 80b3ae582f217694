"""Device-side mirror of the reference's TEST image transform (`get_transform_vit_default(is_train=False)`,
src/pipelines/uni_pipeline.py:1233-1256):

    BGR2RGB -> ToPILImage -> Resize(floor(crop / crop_pct), BICUBIC) -> CenterCrop(crop) -> ToTensor -> Normalize(0.5, 0.5)

The reference runs it per image in the data-loader workers (PIL on the CPU) and uploads fp32 tensors. Here the DECODED 8-bit
images of a batch (cv2 / the TSV decoder deliver HWC BGR uint8 of arbitrary size) are packed into one pinned buffer, uploaded
once and resized + cropped on the GPU (vc_resize_crop_u8, bit-identical to torchvision + Pillow); the result, uint8
(B, crop, crop, 3) in the source channel order, is what `FastImageCaptioning` takes as ``data['image']`` -- the channel flip,
ToTensor and Normalize are fused into its patch extraction (vc_patchify_u8).
"""
import math

import numpy as np
import torch

from . import ops


class DeviceTestTransform:
    """``pixels = DeviceTestTransform(crop_size=384)(list_of_hwc_uint8_images)`` -> uint8 CUDA tensor (B, crop, crop, 3).

    crop_size / crop_pct are the reference's ``test_crop_size`` / ``crop_pct`` (1.0 in the shipped yaml).
    ``test_respect_ratio_max`` (MinMaxResizeForTest) is not on the shipped eval path and is refused."""

    def __init__(self, crop_size=384, crop_pct=1.0, device="cuda", test_respect_ratio_max=None, staging_buffers=2):
        if test_respect_ratio_max:
            raise NotImplementedError("vitcap_b200: MinMaxResizeForTest (test_respect_ratio_max) is not implemented")
        if crop_size % 4:
            raise ValueError("crop_size must be a multiple of 4")
        ops.load_library()
        self.crop = int(crop_size)
        self.resize_to = int(math.floor(crop_size / crop_pct))
        self.device = torch.device(device)
        # pinned staging buffers, rotated: the upload of a batch is asynchronous, so the host may be packing the next batch while
        # the copy of this one is still queued behind earlier GPU work. Each buffer carries the event recorded after its last
        # H2D copy and is rewritten only once that event has completed.
        self._pinned = [None] * max(2, int(staging_buffers))
        self._copied = [None] * len(self._pinned)
        self._slot = 0
        self.h2d_bytes = 0

    def _pack(self, images):
        hw = torch.empty(len(images), 2, dtype=torch.int32)
        off = torch.empty(len(images), dtype=torch.int64)
        total = 0
        arrs = []
        for i, im in enumerate(images):
            a = im.numpy() if torch.is_tensor(im) else np.asarray(im)
            if a.dtype != np.uint8 or a.ndim != 3 or a.shape[2] != 3:
                raise ValueError("image %d: expected uint8 (H, W, 3), got %s %s" % (i, a.dtype, a.shape))
            hw[i, 0], hw[i, 1] = a.shape[0], a.shape[1]
            off[i] = total
            total += (a.size + 15) & ~15                 # 16-byte aligned starts
            arrs.append(a)
        slot = self._slot
        self._slot = (slot + 1) % len(self._pinned)
        if self._copied[slot] is not None:
            self._copied[slot].synchronize()             # the previous upload from this buffer has left the host memory
        if self._pinned[slot] is None or self._pinned[slot].numel() < total:
            self._pinned[slot] = torch.empty(max(total, 1), dtype=torch.uint8).pin_memory()
        flat = self._pinned[slot].numpy()
        for a, o in zip(arrs, off.tolist()):
            flat[o:o + a.size] = a.reshape(-1)
        return self._pinned[slot][:total], off, hw, slot

    def __call__(self, images):
        src_h, off_h, hw_h, slot = self._pack(images)
        B = hw_h.shape[0]
        kmax, max_rows, tmp_off_h = ops.resize_crop_plan(hw_h, self.resize_to, self.crop)
        dev = self.device
        # one small metadata upload: [src_off | tmp_off | hw]
        meta_h = torch.cat([off_h, tmp_off_h[:B], hw_h.view(-1).to(torch.int64)]).pin_memory()
        meta = meta_h.to(dev, non_blocking=True)
        src = src_h.to(dev, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(dev))
        self._copied[slot] = ev
        self.h2d_bytes += src_h.numel() + meta_h.numel() * 8
        hw = meta[2 * B:].to(torch.int32).view(B, 2)
        coef = torch.empty(B * 2 * (kmax + 2) * self.crop, dtype=torch.int32, device=dev)
        tmp = torch.empty(max(int(tmp_off_h[B]), 4), dtype=torch.uint8, device=dev)
        out = torch.empty(B, self.crop, self.crop, 3, dtype=torch.uint8, device=dev)
        return ops.resize_crop_u8(src, meta[:B], hw, self.resize_to, self.crop, kmax, max_rows, coef, tmp, meta[B:2 * B], out)
