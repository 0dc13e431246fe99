"""egn_peer_allreduce (csrc/egn_peer.cu): the one-kernel gradient exchange over peer memory, checked on ONE GPU by running
every "rank" as its own launch on its own stream of the same device -- the handshake protocol, the slice split and the sums
are exactly what N processes on N GPUs execute; only the pointers are local instead of cudaIpc-mapped.  Few blocks per
launch so that all ranks' blocks are co-resident (a block waits for the same block of every peer).  The real N-GPU run is
scripts/peer_check.py (profiles/r02_scaling.md)."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


def _alloc(lib, nbytes):
    from egonerf_b200 import _lib
    p = C.c_void_p()
    _lib.check(lib.egn_peer_alloc(nbytes, C.byref(p)))
    return p.value


@pytest.mark.parametrize("world,n", [(1, 4096), (2, 1 << 20), (4, 1000 * 4), (8, 3 * 4), (3, 123456 * 4)])
def test_peer_allreduce_sums_every_rank_buffer(world, n):
    from egonerf_b200 import _lib
    from egonerf_b200.sharding import _DevicePointer
    lib = _lib.load()
    dev = torch.device("cuda:0")
    bufs = [_alloc(lib, n * 4) for _ in range(world)]
    flags = [_alloc(lib, lib.egn_peer_flag_bytes()) for _ in range(world)]
    holders = [_DevicePointer(b, n) for b in bufs]
    ts = [torch.as_tensor(h, device=dev) for h in holders]
    streams = [torch.cuda.Stream(dev) for _ in range(world)]
    B = (C.c_void_p * world)(*bufs)
    F = (C.c_void_p * world)(*flags)
    g = torch.Generator(device=dev).manual_seed(7)
    try:
        for epoch in (1, 2, 3):                                   # flags are never reset: consecutive calls must work
            src = [torch.randn(n, device=dev, generator=g) for _ in range(world)]
            for t, s in zip(ts, src):
                t.copy_(s)
            want = src[0].clone()
            for s in src[1:]:
                want += s                                          # the kernel's order: rank 0, 1, 2, ...
            want *= 0.5
            torch.cuda.synchronize()
            for r in range(world):
                _lib.check(lib.egn_peer_allreduce(B, F, r, world, n, 0.5, epoch, 8, streams[r].cuda_stream))
            torch.cuda.synchronize()
            for r in range(world):
                assert torch.equal(ts[r], want), f"rank {r}, epoch {epoch}: max diff {(ts[r] - want).abs().max().item()}"
    finally:
        torch.cuda.synchronize()
        del ts
        for p in bufs + flags:
            lib.egn_peer_free(p)


def test_peer_allreduce_rejects_bad_arguments():
    from egonerf_b200 import _lib
    lib = _lib.load()
    b, f = _alloc(lib, 64), _alloc(lib, lib.egn_peer_flag_bytes())
    B, F = (C.c_void_p * 1)(b), (C.c_void_p * 1)(f)
    st = torch.cuda.current_stream().cuda_stream
    try:
        assert lib.egn_peer_allreduce(B, F, 0, 1, 6, 1.0, 1, 8, st) != 0 and b"multiple of 4" in lib.egn_last_error()
        assert lib.egn_peer_allreduce(B, F, 1, 1, 16, 1.0, 1, 8, st) != 0
        assert lib.egn_peer_allreduce(B, F, 0, 1, 16, 1.0, 0, 8, st) != 0 and b"epoch" in lib.egn_last_error()
        assert lib.egn_peer_allreduce(B, F, 0, 1, 16, 1.0, 1, 1000, st) != 0
        assert lib.egn_peer_allreduce(B, F, 0, 17, 16, 1.0, 1, 8, st) != 0
    finally:
        lib.egn_peer_free(b)
        lib.egn_peer_free(f)
