"""CPU, world_size 2 over gloo: the shard / gather plumbing of said_b200.parallel reproduces the
single-process result (the per-rank engine call is replaced by a deterministic stand-in)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from said_b200.parallel import gather_clips, shard_bounds, sharded_inference


class _Stub:
    sampling_rate = 16000

    class denoiser:
        in_channels = 32


def _runner(model, wave, noise, init, mask):
    out = noise * 0.5 + wave.mean(dim=1)[:, None, None]
    if init is not None:
        out = out * (1 - mask) + init * mask
    return out


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(7)
    wave = torch.randn(B, 16000, generator=g)
    init = torch.rand(B, 60, 32, generator=g)
    mask = (torch.rand(B, 60, 32, generator=g) > 0.5).float()
    a = sharded_inference(_Stub, wave, seed=11, runner=_runner)
    b = sharded_inference(_Stub, wave, seed=11, init_samples=init, mask=mask, runner=_runner)
    lo, hi = shard_bounds(B, world, rank)
    c = gather_clips(a[lo:hi].contiguous(), B, dst=0)       # gather on one rank only
    assert (c is None) == (rank != 0)
    if rank == 0:
        assert torch.equal(c, a)
        q.put((a, b))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_bounds_partition():
    for B in (1, 5, 64, 512):
        for w in (1, 2, 4, 8):
            spans = [shard_bounds(B, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
    assert gather_clips(torch.ones(2, 3), 2).shape == (2, 3)    # no process group: identity


@pytest.mark.parametrize("B", [4, 5])
def test_sharded_equals_single_process(B):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, q)) for r in range(2)]
    for p in procs:
        p.start()
    a, b = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = torch.Generator().manual_seed(7)
    wave = torch.randn(B, 16000, generator=g)
    init = torch.rand(B, 60, 32, generator=g)
    mask = (torch.rand(B, 60, 32, generator=g) > 0.5).float()
    gen = torch.Generator().manual_seed(11)
    noise = torch.randn(B, 60, 32, generator=gen)
    assert torch.equal(a, _runner(None, wave, noise, None, None))
    assert torch.equal(b, _runner(None, wave, noise, init, mask))
