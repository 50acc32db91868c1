"""CPU, world_size 2, gloo: the data-parallel shim (nvdiffrecmc_b200/parallel.py).  Sharding the view batch, taking local-mean
losses and averaging ONE flat gradient bucket reproduces the single-process full-batch gradient."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _toy_loss(params, views):
    """Stand-in for 'render views with shared parameters and take the batch-mean image loss': linear-in-light shading
    with a nonlinear material term, per view."""
    light, tex = params
    img = torch.einsum("vp,pc->vpc", views, light) * torch.sigmoid(tex)[None]
    return (img ** 2).mean()


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nvdiffrecmc_b200.parallel import GradBucket, shard_views
    g = torch.Generator().manual_seed(0)
    views = torch.rand(8, 32, generator=g)
    bucket = GradBucket([(32, 3), (32, 3)], device="cpu")
    with torch.no_grad():
        bucket.flat.copy_(torch.rand(bucket.flat.numel(), generator=g))
    sl = shard_views(8)
    assert (sl.stop - sl.start) == 4 and sl.start == rank * 4
    bucket.zero_grad()
    _toy_loss(bucket.params, views[sl]).backward()
    assert bucket.params[0].grad.data_ptr() == bucket.flat_grad.data_ptr()      # autograd accumulated in place into the bucket
    bucket.all_reduce_mean()
    q.put((rank, bucket.flat_grad.clone().numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_bucket_allreduce_equals_full_batch_gradient():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process reference
    from nvdiffrecmc_b200.parallel import GradBucket
    g = torch.Generator().manual_seed(0)
    views = torch.rand(8, 32, generator=g)
    bucket = GradBucket([(32, 3), (32, 3)], device="cpu")
    with torch.no_grad():
        bucket.flat.copy_(torch.rand(bucket.flat.numel(), generator=g))
    _toy_loss(bucket.params, views).backward()
    ref = bucket.flat_grad.numpy()
    assert np.allclose(res[0], res[1], rtol=0, atol=0)
    assert np.allclose(res[0], ref, rtol=1e-5, atol=1e-8)


def test_shard_views_rules():
    from nvdiffrecmc_b200.parallel import shard_views
    assert shard_views(8, rank=3, world=4) == slice(6, 8)
    assert shard_views(8, rank=0, world=1) == slice(0, 8)
    with pytest.raises(ValueError):
        shard_views(6, rank=0, world=4)
