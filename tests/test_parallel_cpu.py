"""world_size-2 gloo tests of the data-parallel host logic (sharding identical to the reference sampler, packed
record all-gather, dataset-order reassembly)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vitcap_b200 import parallel


def test_shard_indices_match_reference_sampler_semantics():
    # contiguous chunks, wrap-around padding (samplers.py:127-146)
    assert parallel.shard_indices(10, 4, 0) == [0, 1, 2]
    assert parallel.shard_indices(10, 4, 3) == [9, 0, 1]
    assert parallel.shard_indices(8, 2, 1) == [4, 5, 6, 7]
    allidx = sum((parallel.shard_indices(11, 3, r) for r in range(3)), [])
    assert allidx[:11] == list(range(11)) and len(allidx) == 12


def test_pack_unpack_roundtrip():
    B, keep, L, K = 5, 2, 20, 50
    ids = torch.randint(0, 30522, (B, keep, L))
    lp = torch.randn(B, keep)
    ti = torch.randint(0, 30522, (B, K))
    tp = torch.rand(B, K)
    rec = parallel.pack_records(ids, lp, ti, tp)
    assert rec.dtype == torch.int32 and rec.shape == (B, keep * L + keep + 2 * K)
    a, b, c, d = parallel.unpack_records(rec, keep, L, K)
    assert torch.equal(a, ids) and torch.equal(b, lp) and torch.equal(c, ti) and torch.equal(d, tp)


class _FakeCaptioner:
    """ids encode the image content so that ordering mistakes are visible. ``per_image`` > 1 mimics num_return_sequences."""

    def __init__(self, per_image=1):
        self.per_image = per_image
        self.last_tags = None

    def __call__(self, data):
        img = data["image"]
        B = img.shape[0]
        tag = img.view(B, -1)[:, 0].long()
        self.last_tags = ((tag.view(B, 1) * 10 + torch.arange(3)).to(torch.int32), tag.float().view(B, 1).repeat(1, 3) / 64)
        tag = tag.repeat_interleave(self.per_image)
        ids = tag.view(-1, 1, 1).repeat(1, 1, 20)
        lp = -tag.float().view(-1, 1)
        return ids, lp


def _worker(rank, world, port, n_items, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        image = torch.arange(n_items).float().view(n_items, 1, 1, 1).repeat(1, 3, 2, 2)
        dp = parallel.DataParallelCaptioner(_FakeCaptioner())
        ids, lp = dp({"image": image, "key": list(range(n_items))})
        # the same with the concept top-k riding along and two returned sequences per image
        dp2 = parallel.DataParallelCaptioner(_FakeCaptioner(per_image=2), with_tags=True)
        ids2, lp2, tidx, tprob = dp2({"image": image, "key": list(range(n_items))})
        want = torch.arange(n_items).repeat_interleave(2)
        assert ids2[:, 0, 0].tolist() == want.tolist() and tuple(ids2.shape) == (2 * n_items, 1, 20)
        assert torch.equal(tidx, want.view(-1, 1) * 10 + torch.arange(3)) and tidx.dtype == torch.int64
        assert torch.equal(tprob, want.float().view(-1, 1).repeat(1, 3) / 64)
        # the side-stream gather degrades to the plain collective where there is no CUDA stream (this test): same rows, rank-major
        rec = torch.full((3, 4), rank, dtype=torch.int32)
        full, done = parallel.SideStreamGather("cpu")(rec)
        assert done is None and full[:, 0].tolist() == sum(([r] * 3 for r in range(world)), [])
        q.put((rank, ids[:, 0, 0].tolist(), lp[:, 0].tolist()))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("n_items", [8, 7])
def test_dp_gather_world2_gloo(n_items):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_items, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ids, lp in res:
        assert ids == list(range(n_items)), (rank, ids)          # every rank holds the whole batch in dataset order
        assert lp == [-float(i) for i in range(n_items)]
