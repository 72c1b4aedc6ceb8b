"""world_size-2 gloo test of the multi-rank host logic: disjoint ray shards per rank, the flat gradient
bucket, one all-reduce(sum) per step, identical parameter updates on every rank.  Runs on CPU — torch tensors only,
no kernels (the CUDA kernels are covered by the -m gpu tests; the NCCL path by bench.py --gpus N)."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _toy_grads(params, x, y):
    """Gradient of a tiny differentiable surrogate (so the test exercises real, shard-dependent gradients)."""
    w3d, w2d, b = params
    pred = (x @ w2d.reshape(w2d.shape[1], -1)[:x.shape[1]]).sum(-1) + w3d.sum() * x.mean(-1) + b.sum()
    loss = ((pred - y) ** 2).sum()          # SUM over the shard: all-reduce(sum) / global count == global mean
    return torch.autograd.grad(loss, params)


def _make_params():
    g = torch.Generator().manual_seed(0)
    w3d = torch.randn(1, 4, 3, 3, 3, generator=g).contiguous(memory_format=torch.channels_last_3d).requires_grad_()
    w2d = torch.randn(1, 6, 5, 2, generator=g).contiguous(memory_format=torch.channels_last).requires_grad_()
    b = torch.randn(7, generator=g).requires_grad_()
    return [w3d, w2d, b]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import ffb200  # noqa: F401
    from ffb200.train import GradBucket, shard_slice
    params = _make_params()
    g = torch.Generator().manual_seed(1)
    N = 64
    x, y = torch.randn(N, 6, generator=g), torch.randn(N, generator=g)     # identical "global batch" on every rank
    sl = shard_slice(N, rank, world)
    grads = _toy_grads(params, x[sl], y[sl])
    bucket = GradBucket(params)
    bucket.pack(grads)
    n = bucket.all_reduce()
    assert n == world
    # views keep each parameter's own strides (channels-last grids), so Adam can run on raw storage
    for i, p in enumerate(params):
        assert bucket.view(i).stride() == p.stride() and bucket.view(i).shape == p.shape
    np.save(os.path.join(out_dir, f'flat_{rank}.npy'), bucket.flat.numpy())
    np.save(os.path.join(out_dir, f'slice_{rank}.npy'), np.array([sl.start, sl.stop]))
    # a None gradient (frozen / unused parameter) must be packed as zeros
    bucket.pack([None] + list(grads[1:]))
    assert float(bucket.view(0).abs().sum()) == 0.0
    dist.barrier()
    dist.destroy_process_group()


def test_grad_bucket_allreduce_world2(tmp_path):
    world, port = 2, _free_port()
    mp.start_processes(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True, start_method='spawn')
    flats = [np.load(tmp_path / f'flat_{r}.npy') for r in range(world)]
    assert np.array_equal(flats[0], flats[1]), 'ranks disagree after the all-reduce'
    # shards are disjoint and cover the batch
    sl = [np.load(tmp_path / f'slice_{r}.npy') for r in range(world)]
    assert sl[0][0] == 0 and sl[0][1] == sl[1][0] and sl[1][1] == 64
    # equals the single-process gradient of the whole batch (sum of shard gradients)
    params = _make_params()
    g = torch.Generator().manual_seed(1)
    x, y = torch.randn(64, 6, generator=g), torch.randn(64, generator=g)
    grads = _toy_grads(params, x, y)
    from ffb200.train import GradBucket
    ref = GradBucket(params)
    ref.pack(grads)
    np.testing.assert_allclose(flats[0], ref.flat.numpy(), rtol=1e-5, atol=1e-5)


def test_shard_slice_covers_everything():
    from ffb200.train import shard_slice
    for n in (0, 1, 7, 64, 4096, 4097):
        for world in (1, 2, 3, 8):
            got = []
            for r in range(world):
                s = shard_slice(n, r, world)
                got.extend(range(s.start, s.stop))
            assert got == list(range(n))


def _worker_render(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import ffb200  # noqa: F401
    from ffb200.train import gather_rows, render_sharded, shard_slice
    N = 1001                                    # odd: the shards differ by one row
    rays = torch.arange(N * 6, dtype=torch.float32).reshape(N, 6)
    fake = lambda r: (r[:, :3] * 2.0 + 1.0, r[:, 3] - r[:, 0])      # stands in for render_ray on this rank's rows
    rgb, depth = render_sharded(rays, None, render_fn=fake)
    np.save(os.path.join(out_dir, f'rgb_{rank}.npy'), rgb.numpy())
    np.save(os.path.join(out_dir, f'depth_{rank}.npy'), depth.numpy())
    sl = shard_slice(N, rank, world)
    assert torch.equal(gather_rows(rays[sl], N), rays)
    dist.barrier()
    dist.destroy_process_group()


def test_render_sharded_world2(tmp_path):
    """Evaluation sharded by ray: every rank ends up with the full maps, equal to the unsharded render."""
    world, port = 2, _free_port()
    mp.start_processes(_worker_render, args=(world, port, str(tmp_path)), nprocs=world, join=True, start_method='spawn')
    rays = torch.arange(1001 * 6, dtype=torch.float32).reshape(1001, 6)
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f'rgb_{r}.npy'), (rays[:, :3] * 2.0 + 1.0).numpy())
        assert np.array_equal(np.load(tmp_path / f'depth_{r}.npy'), (rays[:, 3] - rays[:, 0]).numpy())


def test_image_set_shard_and_arena_ranges():
    """Host logic of image-set training sharded by image: image ranges partition the set, and the all-reduce
    covers exactly the arena slices of the replicated parameters (the coefficient slabs stay local)."""
    from ffb200.train import GradBucket, arena_ranges, image_set_shard
    for n_img, world in [(800, 8), (6, 2), (7, 3)]:
        got = []
        for r in range(world):
            i0, n = image_set_shard(n_img, r, world)
            got.extend(range(i0, i0 + n))
        assert got == list(range(n_img))
    params = [torch.zeros(64, 36), torch.zeros(64), torch.zeros(3, 64), torch.zeros(1, 36, 3, 27, 27), torch.zeros(1, 8, 8, 8), torch.zeros(1, 4, 13, 13)]
    bucket = GradBucket(params)
    ranges = arena_ranges(bucket, [True, True, True, False, True, True])        # get_optparam_groups order: MLP, coeffs, bases
    assert len(ranges) == 2 and ranges[0][0] == 0 and ranges[0][1] == bucket.offsets[3] and ranges[1] == (bucket.offsets[4], bucket.offsets[6])
    covered = sum(b - a for a, b in ranges)
    assert covered == bucket.offsets[-1] - (bucket.offsets[4] - bucket.offsets[3])
    assert arena_ranges(bucket, [True] * 6) == [(0, bucket.offsets[-1])]
