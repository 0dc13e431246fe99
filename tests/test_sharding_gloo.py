"""CPU, world_size 2 over gloo: the host-side logic of ray-sharded data parallelism — shard ranges,
the flat gradient bucket and its single all-reduce.  The device side (per-ray RNG keyed by the global ray index, shard
additivity of gradients) is covered on the GPU by tests/test_gpu_grad.py."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from egonerf_b200.sharding import GradientBucket, row_tiles, shard_range


def test_shard_ranges_partition_exactly():
    for n in (0, 1, 7, 4096, 65536, 131072, 4096 * 2048 + 3):
        for world in (1, 2, 3, 4, 8):
            blocks = [shard_range(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1
    assert row_tiles(2048, 3, 8) == (768, 1024)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        params = [torch.nn.Parameter(torch.randn(s)) for s in ((1, 16, 22, 20), (1, 48, 64, 1), (27, 144), (128,), (3, 8, 4))]
        bucket = GradientBucket(params)
        # every rank computes the gradient of its own ray shard of a toy "render": loss = sum_r w_r * <p, x_r>
        n = 1001
        a, b = shard_range(n, rank, world)
        g = torch.Generator().manual_seed(5)
        coeff = torch.rand(n, generator=g)
        for p in params:
            p.grad = None
        loss = sum((p * p).sum() for p in params) * coeff[a:b].sum()
        loss.backward()                                   # autograd allocates its own .grad tensors
        bucket.gather_from_params()
        assert all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(params, bucket.views))
        bucket.allreduce()
        expect = [2 * p.detach() * coeff.sum() for p in params]
        err = max(float((p.grad - e).abs().max() / e.abs().max()) for p, e in zip(params, expect))
        out[rank] = err
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_gradient_bucket_allreduce_world2():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert len(out) == world and all(e < 1e-5 for e in out.values()), dict(out)


def test_peer_slices_partition_the_buffer():
    """Slice split of egn_peer_allreduce (float4 units, ceil split): slices are disjoint, ordered and cover the buffer."""
    from egonerf_b200.sharding import slice_bounds
    for n in (4, 8, 12, 4000, 24721124 + 56320, 104):
        for world in (1, 2, 3, 4, 8, 16):
            prev = 0
            for r in range(world):
                lo, hi = slice_bounds(n, r, world)
                assert lo == prev and lo % 4 == 0 and hi % 4 == 0 and hi >= lo
                prev = hi
            assert prev == n


def test_peer_exchange_is_not_used_without_a_process_group():
    """enable_peer_exchange is a no-op (False) outside a multi-rank process group: single-GPU training is untouched."""
    import types
    from egonerf_b200.optim import TableAdam
    opt = TableAdam.__new__(TableAdam)
    opt.peer = None
    opt.model = types.SimpleNamespace()
    assert TableAdam.enable_peer_exchange(opt) is False and opt.peer is None


def test_gradient_bucket_on_external_storage():
    """The exchange bucket may live in someone else's buffer (the tail of the peer-memory gradient buffer): the views alias
    that storage, zero() clears exactly the bucket's part and re-attaches, gather_from_params copies only foreign gradients."""
    import torch
    from egonerf_b200.sharding import GradientBucket
    a, b = torch.nn.Parameter(torch.zeros(3, 2)), torch.nn.Parameter(torch.zeros(5))
    store = torch.full((16,), 7.0)
    bucket = GradientBucket([a, b], storage=store)
    assert bucket.external and bucket.flat.data_ptr() == store.data_ptr() and bucket.flat.numel() == 11
    assert float(store[:11].abs().sum()) == 0.0 and float(store[11:].sum()) == 5 * 7.0      # only the bucket's part was cleared
    a.grad = torch.ones(3, 2)                     # a gradient autograd allocated on its own
    bucket.gather_from_params()
    assert a.grad.data_ptr() == bucket.views[0].data_ptr() and b.grad.data_ptr() == bucket.views[1].data_ptr()
    assert torch.equal(store[:6], torch.ones(6)) and float(store[6:11].abs().sum()) == 0.0
    b.grad.add_(2.0)                              # in-place accumulation lands in the external storage
    assert torch.equal(store[6:11], torch.full((5,), 2.0))
    bucket.zero()
    assert float(store[:11].abs().sum()) == 0.0 and a.grad.data_ptr() == store.data_ptr()
