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


def _worker_f1(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nvdiffrecmc_b200.parallel import GradBucket, StepRNG, hook_optimizer, shard_indices, shard_views, sync_seed
    np.random.seed(100 + rank)                                  # ranks start with DIFFERENT process-global RNG states
    seed = sync_seed()
    rng = StepRNG(seed)
    jitter = torch.normal(0.0, 1.0, (4,), generator=rng.for_step(7))
    idx = shard_indices(n_items=10, global_batch=4, it=3, seed=seed)
    sl = shard_views(8)
    per_view = torch.stack([torch.rand(3, generator=rng.for_view(7, v)) for v in range(sl.start, sl.stop)])
    # two optimizer steps of the toy problem with the hooked optimizer
    g = torch.Generator().manual_seed(0)
    views = torch.rand(8, 32, generator=g)
    bucket = GradBucket([(32, 3), (32, 3)], device="cpu")
    with torch.no_grad():
        bucket.flat.copy_(torch.rand(bucket.flat.numel(), generator=g))
    opt = hook_optimizer(torch.optim.Adam(bucket.params, lr=0.05), bucket)
    for _ in range(2):
        bucket.zero_grad()
        _toy_loss(bucket.params, views[sl]).backward()
        opt.step()
    q.put((rank, seed, jitter.numpy(), idx.numpy(), per_view.numpy(), bucket.flat.detach().clone().numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_training_loop_shim_rank_consistency():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_f1, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in range(2)), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, s0, j0, i0, v0, f0), (_, s1, j1, i1, v1, f1) = res
    assert s0 == s1 and np.array_equal(j0, j1)                                    # one seed, one per-step stream
    from nvdiffrecmc_b200.parallel import GradBucket, StepRNG, shard_indices
    full = shard_indices(10, 4, 3, s0, rank=0, world=1).numpy()
    assert np.array_equal(np.concatenate([i0, i1]), full) and len(set(full.tolist())) == 4       # shards tile the single-process batch
    rng = StepRNG(s0)
    assert np.array_equal(np.concatenate([v0, v1]), torch.stack([torch.rand(3, generator=rng.for_view(7, v)) for v in range(8)]).numpy())
    assert not np.array_equal(v0[0], v0[1])
    # parameters after two hooked Adam steps == single process on the full batch
    g = torch.Generator().manual_seed(0)
    views = torch.rand(8, 32, generator=g)
    bucket = GradBucket([(32, 3), (32, 3)], device="cpu")
    with torch.no_grad():
        bucket.flat.copy_(torch.rand(bucket.flat.numel(), generator=g))
    opt = torch.optim.Adam(bucket.params, lr=0.05)
    for _ in range(2):
        bucket.zero_grad()
        _toy_loss(bucket.params, views).backward()
        opt.step()
    assert np.array_equal(f0, f1) and np.allclose(f0, bucket.flat.detach().numpy(), rtol=1e-5, atol=1e-7)


def test_shard_indices_epochs():
    from nvdiffrecmc_b200.parallel import shard_indices
    seen = np.concatenate([shard_indices(12, 4, it, seed=5, rank=0, world=1).numpy() for it in range(3)])
    assert sorted(seen.tolist()) == list(range(12))                               # one epoch visits every item once
    nxt = np.concatenate([shard_indices(12, 4, it, seed=5, rank=0, world=1).numpy() for it in range(3, 6)])
    assert sorted(nxt.tolist()) == list(range(12)) and not np.array_equal(nxt, seen)      # the next epoch is a new permutation
    assert shard_indices(3, 4, 0, seed=1, rank=0, world=2).numel() == 2           # dataset smaller than the batch wraps around


# ---- the reference loop calls optimizer.zero_grad() (set_to_none=True) and builds a LambdaLR per optimizer ------------
def _worker_ref_loop(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nvdiffrecmc_b200.parallel import GradBucket, hook_optimizer, shard_views
    g = torch.Generator().manual_seed(0)
    views = torch.rand(8, 32, generator=g)
    # parameters owned by "modules", as in train.py:340-356; the bucket adopts them in place
    light = torch.nn.Parameter(torch.rand(32, 3, generator=g)); tex = torch.nn.Parameter(torch.rand(32, 3, generator=g))
    bucket = GradBucket.adopt([light, tex])
    assert light.data_ptr() == bucket.flat.data_ptr()
    opt = torch.optim.Adam([light, tex], lr=0.05)
    order = rank % 2 == 0                                        # scheduler before / after the hook: both must work
    if order:
        sched = torch.optim.lr_scheduler.LambdaLR(opt, lr_lambda=lambda it: 0.5 ** it)
    hook_optimizer(opt, bucket)
    if not order:
        sched = torch.optim.lr_scheduler.LambdaLR(opt, lr_lambda=lambda it: 0.5 ** it)
    sl = shard_views(8)
    for _ in range(3):
        opt.zero_grad()                                          # train.py:407-411 (set_to_none=True by default)
        _toy_loss((light, tex), views[sl]).backward()
        opt.step()                                               # train.py:452
        sched.step()                                             # train.py:453
    q.put((rank, bucket.flat.detach().clone().numpy(), float(sched.get_last_lr()[0])))
    dist.barrier()
    dist.destroy_process_group()


def test_reference_loop_shape_zero_grad_and_scheduler():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_ref_loop, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in range(2)), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = torch.Generator().manual_seed(0)
    views = torch.rand(8, 32, generator=g)
    light = torch.nn.Parameter(torch.rand(32, 3, generator=g)); tex = torch.nn.Parameter(torch.rand(32, 3, generator=g))
    opt = torch.optim.Adam([light, tex], lr=0.05)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lr_lambda=lambda it: 0.5 ** it)
    for _ in range(3):
        opt.zero_grad()
        _toy_loss((light, tex), views).backward()
        opt.step(); sched.step()
    ref = torch.cat([light.detach().reshape(-1), tex.detach().reshape(-1)]).numpy()
    assert np.array_equal(res[0][1], res[1][1]), "ranks diverged"
    assert np.allclose(res[0][1], ref, rtol=1e-5, atol=1e-7)
    assert res[0][2] == res[1][2] == float(sched.get_last_lr()[0])


def test_bucket_survives_set_to_none_without_the_hook():
    """all_reduce_mean() must never reduce a stale bucket: foreign .grad tensors are copied in, missing ones zero their slice."""
    from nvdiffrecmc_b200.parallel import GradBucket
    b = GradBucket([(3,), (2,)], device="cpu")
    opt = torch.optim.SGD(b.params, lr=0.1)
    b.flat_grad.fill_(7.0)                                       # stale content from an earlier step
    opt.zero_grad()                                              # drops the aliases
    (b.params[0] * 2).sum().backward()                           # fresh .grad for params[0] only
    assert b.params[0].grad.data_ptr() != b.flat_grad.data_ptr()
    b.all_reduce_mean()
    assert b.flat_grad.tolist() == [2.0, 2.0, 2.0, 0.0, 0.0]
    assert b.params[0].grad.data_ptr() == b.flat_grad.data_ptr() and b.params[1].grad is not None
