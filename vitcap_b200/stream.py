"""Host-side feeding loop of the caption path: the reference's ``predict_iter`` (uni_pipeline.py:709-730) moves every
batch to the GPU (``recursive_to_device``) and then runs the model, one after the other on one stream. At the B200's
caption rate the 906 MB of fp32 pixels per 512-image batch cost as much PCIe time as 14 % of the compute, so the drop-in
loop here uploads batch i+1 on a copy stream while batch i is being captioned, and reads the small result records back
into pinned memory. Same inputs (pinned host tensors with the reference's dict keys), same outputs, same order.
"""
from collections import deque

import torch

from . import parallel


class OverlappedCaptioner:
    """``for ids, logprobs in OverlappedCaptioner(model).run(host_batches)`` -- ``host_batches`` yields dicts of (ideally
    pinned) CPU tensors as produced by the reference's data loader; results are CPU tensors (ids int64 (B, keep, L),
    logprobs fp32 (B, keep)) of the whole batch gathered over all ranks when torch.distributed is initialised."""

    def __init__(self, model, device=None, depth=2, gather=True, with_tags=False):
        self.model = model
        self.device = torch.device(device) if device is not None else next(model.parameters()).device
        self.depth = depth
        self.gather = gather
        self.with_tags = with_tags
        self.copy_stream = torch.cuda.Stream(device=self.device)
        # the cross-rank gather and the read-back of batch i run on a side stream while the compute stream starts batch i+1
        self.side_gather = parallel.SideStreamGather(self.device) if gather else None
        self._dev = [dict() for _ in range(depth)]       # per-slot device staging buffers
        self._free = [None] * depth                      # event: compute on the slot's buffers has finished
        self._out = [None] * depth                       # per-slot pinned result buffer
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def _upload(self, slot, hb):
        cs = self.copy_stream
        if self._free[slot] is not None:
            cs.wait_event(self._free[slot])
        bufs = self._dev[slot]
        out = {}
        with torch.cuda.stream(cs):
            for k, v in hb.items():
                if not torch.is_tensor(v):
                    out[k] = v
                    continue
                b = bufs.get(k)
                if b is None or b.shape != v.shape or b.dtype != v.dtype:
                    b = torch.empty(v.shape, dtype=v.dtype, device=self.device)
                    bufs[k] = b
                b.copy_(v, non_blocking=True)
                self.h2d_bytes += v.numel() * v.element_size()
                out[k] = b
            ready = torch.cuda.Event()
            ready.record(cs)
        return out, ready

    def _caption(self, slot, dev_batch, ready):
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ready)
        ids, lp = self.model(dev_batch)
        keep, max_len = ids.shape[1], ids.shape[2]
        if self.with_tags:
            tag_idx, tag_prob = self.model.last_tags
            per_image = ids.shape[0] // tag_idx.shape[0]          # num_return_sequences rows per image
            if per_image > 1:
                tag_idx, tag_prob = tag_idx.repeat_interleave(per_image, 0), tag_prob.repeat_interleave(per_image, 0)
            rec = parallel.pack_records(ids, lp, tag_idx, tag_prob)
        else:
            rec = parallel.pack_records(ids, lp)
        free = torch.cuda.Event()                        # the slot's device staging buffers may be refilled after this point
        free.record(cur)
        self._free[slot] = free
        side = self.side_gather is not None and self.side_gather.active()
        if side:
            full, gathered = self.side_gather(rec)
            out_stream = self.side_gather.stream
        else:
            full, out_stream = rec, cur
        n_rows = full.shape[0]
        o = self._out[slot]
        if o is None or o.shape != full.shape:
            o = torch.empty(full.shape, dtype=full.dtype).pin_memory()
            self._out[slot] = o
        with torch.cuda.stream(out_stream):
            o.copy_(full, non_blocking=True)
            done = torch.cuda.Event()
            done.record(out_stream)
        self.d2h_bytes += n_rows * full.shape[1] * full.element_size()
        return o, done, keep, max_len

    def run(self, host_batches):
        """Software pipeline of depth ``self.depth``: batch i is uploaded AND its kernels are queued as soon as the host hands
        it over; the host then waits for the oldest batch in flight. The ~170 eager launches of batch i+1 (a few ms of host
        time) are issued while batch i is still running, so the GPU never waits for the host between batches."""
        inflight = deque()
        i = 0
        for hb in host_batches:
            slot = i % self.depth
            i += 1
            if len(inflight) == self.depth:              # the slot's staging and result buffers are about to be reused
                yield self._finish(*inflight.popleft())
            dev_batch, ready = self._upload(slot, hb)
            inflight.append(self._caption(slot, dev_batch, ready))
        while inflight:
            yield self._finish(*inflight.popleft())

    def _finish(self, o, done, keep, max_len):
        done.synchronize()                               # the caller reads the captions on the host
        topk = self.model.cfg.topk if self.with_tags else None
        return parallel.unpack_records(o.clone(), keep, max_len, topk)
