"""Multi-GPU: shard independent clips across ranks, gather results once.

Clips are independent units (no BatchNorm anywhere on the path; GroupNorm / LayerNorm / attention /
``rescale_noise_cfg`` are all per clip), so there is no data-path collective: rank ``r`` of ``N`` runs
the whole pipeline on clips ``[r*B/N, (r+1)*B/N)`` and the only communication is one all-gather of the
``(B/N, T, 32)`` results.  One process per GPU (``torchrun``); ``torch.distributed`` supplies the NCCL
communicator over NVLink.  On CPU test rigs the same code runs over ``gloo`` with a stand-in per-rank
worker, which is how the shard / gather index math is tested without GPUs.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(batch: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced partition: the first ``batch % world_size`` ranks get one extra clip."""
    base, extra = divmod(batch, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def draw_sharded_noise(shape, seed: int, device, lo: int, hi: int) -> torch.Tensor:
    """Every rank draws the FULL batch's noise from the same seed and keeps its slice, so the union over
    ranks is bit-identical to the single-device run with ``torch.manual_seed(seed)`` -- for a CUDA
    ``device`` that is the draw ``SAID.inference`` makes (``diffusion.py:364``)."""
    gen = torch.Generator(device=device)
    gen.manual_seed(int(seed))
    full = torch.randn(shape, device=device, generator=gen)
    return full[lo:hi].contiguous()


def gather_clips(local: torch.Tensor, batch: int, group=None, dst: Optional[int] = None) -> Optional[torch.Tensor]:
    """Gather ragged per-rank results ``(b_r, ...)`` into ``(batch, ...)``: on every rank (all-gather, ``dst=None``) or
    on rank ``dst`` only (the other ranks return ``None``) -- a caller that writes the CSVs on one rank needs no more."""
    if not (dist.is_available() and dist.is_initialized()):
        return local
    world = dist.get_world_size(group)
    if world == 1:
        return local
    rank = dist.get_rank(group)
    sizes = [shard_bounds(batch, world, r) for r in range(world)]
    max_b = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((max_b,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    if dst is None:
        bufs = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(bufs, pad, group=group)
    else:
        bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
        dist.gather(pad, bufs, dst=dst, group=group)
        if rank != dst:
            return None
    return torch.cat([bufs[r][: hi - lo] for r, (lo, hi) in enumerate(sizes)], dim=0)


def sharded_inference(
    model,
    waveform_processed: torch.Tensor,
    seed: int,
    init_samples: Optional[torch.Tensor] = None,
    mask: Optional[torch.Tensor] = None,
    num_inference_steps: int = 100,
    strength: float = 1.0,
    guidance_scale: float = 2.5,
    guidance_rescale: float = 0.0,
    fps: int = 60,
    group=None,
    gather: bool = True,
    runner: Optional[Callable] = None,
):
    """``model.inference`` over a batch sharded across the process group.

    ``waveform_processed`` (and ``init_samples`` / ``mask``) hold the FULL batch on every rank (host or
    device); each rank moves only its slice to its GPU.  Deterministic (eta = 0) sampling only.  Returns
    the gathered ``(B, T, C)`` result on every rank (``gather=False``: the local slice).
    ``runner(model, wave, noise, init, mask) -> Tensor`` replaces the engine call in CPU tests.
    """
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    B, T_a = waveform_processed.shape
    lo, hi = shard_bounds(B, world, rank)
    device = next(model.parameters()).device if runner is None else waveform_processed.device
    T = int(T_a / model.sampling_rate * fps)
    C = model.denoiser.in_channels
    noise = draw_sharded_noise((B, T, C), seed, device, lo, hi)
    wave = waveform_processed[lo:hi].to(device)
    init = None if init_samples is None else init_samples[lo:hi].to(device)
    msk = None if mask is None else mask[lo:hi].to(device)
    if hi > lo:
        if runner is not None:
            local = runner(model, wave, noise, init, msk)
        else:
            local = model._run(wave, noise, init, msk, num_inference_steps, strength, guidance_scale,
                               guidance_rescale, 0.0, T, False, False, None).result
    else:
        local = torch.empty((0, T, C), dtype=torch.float32, device=device)
    return gather_clips(local, B, group) if gather else local
