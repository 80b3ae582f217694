"""Data-parallel plumbing: one process per GPU, images sharded in contiguous chunks, ONE collective per batch.

The reference shards the test set with a contiguous-chunk ``DistributedSampler`` (src/data_layer/samplers.py:86-146:
pad by wrapping, rank r takes [r*n, (r+1)*n)), has every rank write ``..._{rank}_{size}.tsv`` and lets rank 0
concatenate / re-order / de-duplicate the files behind a barrier (uni_pipeline.py:783-831). Here the same sharding
is kept and the files are replaced by a single all-gather of a packed per-image record
{ids int32[max_len*keep], logprob f32[keep], tag_idx int32[K], tag_prob f32[K]} (484 B per image at the defaults)
over NCCL (NVLink 5 / NVSwitch); gloo is used by the CPU tests.
"""
import math

import torch
import torch.distributed as dist


def bind_to_gpu_numa(device_index):
    """Pins this process (and the pinned host buffers it allocates afterwards) to the CPU cores NVML reports as local to GPU
    ``device_index``: with one process per GPU on a two-socket host, uploads of 0.9 GB per batch otherwise cross the
    inter-socket link for half of the ranks. Returns the core list, or None when NVML / the affinity call is unavailable
    (nothing is changed then). Call it before allocating pinned memory; undo with os.sched_setaffinity(0, previous)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = device_index
        if vis:
            ent = vis.split(",")[device_index].strip()
            if ent.isdigit():
                phys = int(ent)
            else:
                return None
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cores = [w * 64 + b for w, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cores = sorted(set(cores) & set(allowed))
        if not cores:
            return None
        os.sched_setaffinity(0, cores)
        return cores
    except Exception:
        return None


def shard_indices(n_items, world_size, rank):
    """Indices of the reference's non-shuffled DistributedSampler for ``rank`` (samplers.py:127-146)."""
    num_samples = int(math.ceil(n_items * 1.0 / world_size))
    total = num_samples * world_size
    idx = list(range(n_items))
    assert total - len(idx) <= len(idx), "not implemented (same limit as the reference)"
    idx += idx[: total - len(idx)]
    return idx[num_samples * rank: num_samples * (rank + 1)]


def pack_records(ids, logprobs, tag_idx=None, tag_prob=None):
    """(B,keep,L) int64, (B,keep) f32, (B,K) int, (B,K) f32 -> int32 (B, W) record matrix (floats bit-cast)."""
    B = ids.shape[0]
    parts = [ids.reshape(B, -1).to(torch.int32), logprobs.reshape(B, -1).float().contiguous().view(torch.int32)]
    if tag_idx is not None:
        parts.append(tag_idx.reshape(B, -1).to(torch.int32))
        parts.append(tag_prob.reshape(B, -1).float().contiguous().view(torch.int32))
    return torch.cat(parts, dim=1).contiguous()


def unpack_records(rec, keep, max_len, topk=None):
    B = rec.shape[0]
    o = 0
    ids = rec[:, o:o + keep * max_len].to(torch.int64).view(B, keep, max_len)
    o += keep * max_len
    lp = rec[:, o:o + keep].contiguous().view(torch.float32).view(B, keep)
    o += keep
    if topk is None:
        return ids, lp
    tag_idx = rec[:, o:o + topk].to(torch.int64)
    o += topk
    tag_prob = rec[:, o:o + topk].contiguous().view(torch.float32)
    return ids, lp, tag_idx, tag_prob


def all_gather_records(rec, group=None):
    """rec int32 (b, W) per rank (same b on every rank) -> (world*b, W), rank-major == dataset order of the
    contiguous-chunk sharding."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return rec
    world = dist.get_world_size(group)
    out = torch.empty(world * rec.shape[0], rec.shape[1], dtype=rec.dtype, device=rec.device)
    dist.all_gather_into_tensor(out, rec.contiguous(), group=group)
    return out


class SideStreamGather:
    """all_gather_records off the compute stream. The collective is the path's only cross-rank dependency; issued on the
    compute stream it puts all ranks in lockstep every batch (each step then costs the slowest of N power-capped boards:
    round-1 scaling 0.979 at 8 GPUs). Here the records of batch i are gathered on a dedicated stream -- which waits for the
    event that marks them ready -- while the compute stream already runs batch i+1, so ranks only meet inside the side stream.

        full, done = gather(rec)      # rec was produced on the current stream; `full` is valid once `done` has completed
    """

    def __init__(self, device, group=None):
        self.group = group
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device) if self.device.type == "cuda" else None

    def active(self):
        return self.stream is not None and dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1

    def __call__(self, rec):
        if not self.active():
            return all_gather_records(rec, self.group), None
        cur = torch.cuda.current_stream(self.device)
        ready = torch.cuda.Event()
        ready.record(cur)
        self.stream.wait_event(ready)
        rec.record_stream(self.stream)                   # the caching allocator must not recycle it under the collective
        with torch.cuda.stream(self.stream):
            full = all_gather_records(rec, self.group)
            done = torch.cuda.Event()
            done.record(self.stream)
        return full, done


def gather_in_dataset_order(rec, n_items, group=None):
    """All-gather + drop the wrap-around padding rows (the reference de-duplicates by key on rank 0, uni_pipeline.py:822-828)."""
    full = all_gather_records(rec, group)
    return full[:n_items]


class DataParallelCaptioner:
    """Runs a FastImageCaptioning replica on this rank's shard and returns the gathered results of the whole batch:
    (ids, logprobs), plus (tag_idx, tag_prob) of the concept head when ``with_tags`` is set."""

    def __init__(self, model, group=None, with_tags=False):
        self.model = model
        self.group = group
        self.with_tags = with_tags

    @torch.no_grad()
    def __call__(self, data):
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        rank = dist.get_rank(self.group) if dist.is_initialized() else 0
        n = data["image"].shape[0]
        idx = shard_indices(n, world, rank)
        sel = torch.as_tensor(idx, device=data["image"].device)
        sub = {k: (v.index_select(0, sel) if torch.is_tensor(v) and v.shape[:1] == (n,) else v) for k, v in data.items()}
        if "key" in data and not torch.is_tensor(data["key"]):
            sub["key"] = [data["key"][i] for i in idx]
        ids, lp = self.model(sub)
        keep, max_len = ids.shape[1], ids.shape[2]
        per_image = ids.shape[0] // len(idx)             # num_return_sequences rows per image
        tags = getattr(self.model, "last_tags", None) if self.with_tags else None
        if tags is not None:
            # one record row per returned sequence: repeat the image's concept top-k for each of its sequences
            rec = pack_records(ids, lp, tags[0].repeat_interleave(per_image, 0), tags[1].repeat_interleave(per_image, 0))
        else:
            rec = pack_records(ids, lp)
        full = gather_in_dataset_order(rec, n * per_image, self.group)
        return unpack_records(full, keep, max_len, tags[0].shape[1] if tags is not None else None)
