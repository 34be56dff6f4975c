"""Synthetic weights and audio for tests and benchmarks.

No pretrained SAiD checkpoint or dataset is reachable offline, so parity tests and
``bench.py`` run on seeded synthetic weights in the reference's state-dict layout and on the synthetic
waveform BASELINE.md section 3 defines.  Everything is drawn from an explicitly seeded CPU generator, so the
same tensors are produced in the build container (where golden vectors are made from the reference)
and on the GPU box.

Every tensor is drawn -- including the ones the reference zero-initialises (``zero_module``:
``said/model/ldm/openaimodel.py:182-185, 668``, ``said/model/ldm/attention.py:221``): with those left
at zero the denoiser returns exactly 0 and every comparison would pass vacuously.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch

from .model.params import (
    Wav2Vec2Dims,
    audio_encoder_spec,
    denoiser_spec,
)


def _draw(name: str, shape, gen: torch.Generator) -> torch.Tensor:
    """One tensor, scaled by role so activations stay O(1) through the networks."""
    leaf = name.rsplit(".", 1)[-1]
    is_norm = (
        ".layer_norm." in name
        or "final_layer_norm" in name
        or ".norm" in name
        or name.startswith("out.0.")
        or ".in_layers.0." in name
        or ".out_layers.0." in name
    )
    if name.endswith("weight_g"):
        # weight-norm magnitude of the positional conv: keep the folded kernel near unit gain
        return 1.0 + 0.25 * torch.randn(shape, generator=gen)
    if is_norm:
        if leaf == "weight":
            return 1.0 + 0.1 * torch.randn(shape, generator=gen)
        return 0.1 * torch.randn(shape, generator=gen)
    if leaf == "bias":
        return 0.05 * torch.randn(shape, generator=gen)
    if name == "masked_spec_embed":
        return torch.rand(shape, generator=gen)
    fan_in = 1
    for s in shape[1:]:
        fan_in *= s
    if "feature_extractor.conv_layers" in name:
        std = math.sqrt(2.0 / fan_in)  # GELU stack without normalisation after layer 0
    else:
        std = 1.0 / math.sqrt(fan_in)
    return std * torch.randn(shape, generator=gen)


def synthetic_state_dict(
    seed: int = 0,
    in_channels: int = 32,
    feature_dim: int = -1,
    audio_dims: Optional[Wav2Vec2Dims] = None,
    dtype: torch.dtype = torch.float32,
) -> Dict[str, torch.Tensor]:
    """State dict with the reference's 372 keys (373-374 with ``feature_dim > 0``)."""
    d = audio_dims or Wav2Vec2Dims()
    ctx = feature_dim if feature_dim > 0 else d.hidden
    gen = torch.Generator(device="cpu").manual_seed(int(seed))
    sd: Dict[str, torch.Tensor] = {}
    for k, shape in audio_encoder_spec(d):
        sd["audio_encoder." + k] = _draw(k, shape, gen).to(dtype)
    for k, shape in denoiser_spec(in_channels, ctx):
        sd["denoiser.model." + k] = _draw(k, shape, gen).to(dtype)
    sd["null_cond_emb"] = torch.randn((1, 1, ctx), generator=gen).to(dtype)
    if feature_dim > 0:
        sd["audio_proj_layer.weight"] = (
            torch.randn((feature_dim, d.output_hidden), generator=gen) / math.sqrt(d.output_hidden)
        ).to(dtype)
        sd["audio_proj_layer.bias"] = (0.05 * torch.randn((feature_dim,), generator=gen)).to(dtype)
    return sd


def synthetic_waveform(clip_index: int = 0, seconds: float = 5.0, sampling_rate: int = 16000) -> np.ndarray:
    """``0.3 sin(2 pi f1 t) + 0.1 sin(2 pi 3 t) sin(2 pi f2 t)``, f1 = 220(1+i/64), f2 = 4 f1."""
    n = int(round(seconds * sampling_rate))
    t = np.arange(n, dtype=np.float64) / sampling_rate
    f1 = 220.0 * (1.0 + clip_index / 64.0)
    f2 = 880.0 * (1.0 + clip_index / 64.0)
    x = 0.3 * np.sin(2 * np.pi * f1 * t) + 0.1 * np.sin(2 * np.pi * 3.0 * t) * np.sin(2 * np.pi * f2 * t)
    return x.astype(np.float32)


def normalise_waveform(x: np.ndarray) -> np.ndarray:
    """Per-utterance zero-mean / unit-variance (HF ``Wav2Vec2FeatureExtractor.zero_mean_unit_var_norm``,
    ``feature_extraction_wav2vec2.py:78-97``): ``(x - mean) / sqrt(var + 1e-7)``, population variance."""
    x = np.asarray(x, dtype=np.float32)
    return ((x - x.mean()) / np.sqrt(x.var() + 1e-7)).astype(np.float32)


def synthetic_batch(batch: int, seconds: float = 5.0, sampling_rate: int = 16000) -> torch.Tensor:
    """``(batch, T_a)`` processed (normalised) waveforms, one distinct clip per row."""
    rows = [normalise_waveform(synthetic_waveform(i, seconds, sampling_rate)) for i in range(batch)]
    return torch.from_numpy(np.stack(rows, 0))
