"""CPU oracle (TEST INFRASTRUCTURE, never imported by the product path) for the distribution-level evaluation that follows the
hot path: the BCVAE encoder over sliding 120-frame windows and the Frechet distance between two sets of
latents.  Restates

* ``said/model/vae.py:26-112``            ``BCEncoder`` in eval mode (BatchNorm with running statistics) -> ``mean``;
* ``script/test_evaluate.py:53-106``      the windowing: ``num_windows = (len - 120) // step + 1``, window = frames [s, s + 120);
* ``said/metric/frechet_distance.py:17-64`` ``get_statistic`` (mean, ``np.cov(rowvar=False)``) and ``frechet_distance``, which calls
  ``pytorch_fid.fid_score.calculate_frechet_distance`` (third-party, pinned by the reference's ``pyproject.toml``, absent from
  the container): ``|mu1 - mu2|^2 + Tr(S1) + Tr(S2) - 2 Tr(sqrtm(S1 S2))`` with ``scipy.linalg.sqrtm`` -- restated from the
  published algorithm, PARITY UNPINNED at that boundary (no golden vectors exist for it; the encoder half is pinned against the
  reference's own ``BCVAE`` class by ``tests/golden/bcvae_windows.npz``).
"""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]
SEQ_LEN = 120


def bcvae_encoder_spec(in_channels: int = 32, z_dim: int = 64):
    """Names / shapes of ``BCVAE.encoder.*`` (``vae.py:41-66``) that the eval-mode forward reads."""
    spec = []

    def conv(i, co, ci, k):
        spec.append((f"encoder.conv_layers.{i}.weight", (co, ci, k)))
        spec.append((f"encoder.conv_layers.{i}.bias", (co,)))

    def bn(prefix, c):
        for leaf in ("weight", "bias", "running_mean", "running_var"):
            spec.append((f"{prefix}.{leaf}", (c,)))

    conv(0, 32, in_channels, 3); bn("encoder.conv_layers.1", 32)
    conv(3, 64, 32, 3); bn("encoder.conv_layers.4", 64)
    conv(6, 64, 64, 4); bn("encoder.conv_layers.7", 64)
    conv(9, 32, 64, 3)
    spec += [("encoder.fc_layers.0.weight", (256, 1760)), ("encoder.fc_layers.0.bias", (256,))]
    bn("encoder.fc_layers.1", 256)
    spec += [("encoder.fc_layers.3.weight", (128, 256)), ("encoder.fc_layers.3.bias", (128,))]
    bn("encoder.fc_layers.4", 128)
    spec += [("encoder.fc_layers.6.weight", (z_dim, 128)), ("encoder.fc_layers.6.bias", (z_dim,)),
             ("encoder.fc_mu.weight", (z_dim, z_dim)), ("encoder.fc_mu.bias", (z_dim,))]
    return spec


def synthetic_bcvae_state_dict(seed: int = 0) -> SD:
    """Seeded stand-in for ``model/vae.pth`` (the trained file stays with the reference): encoder tensors only."""
    g = torch.Generator().manual_seed(seed)
    sd: SD = {}
    for name, shape in bcvae_encoder_spec():
        leaf = name.rsplit(".", 1)[-1]
        if leaf == "running_var":
            sd[name] = 0.5 + torch.rand(shape, generator=g)
        elif leaf == "running_mean":
            sd[name] = 0.1 * torch.randn(shape, generator=g)
        elif len(shape) == 1 and leaf == "weight":
            sd[name] = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif leaf == "bias":
            sd[name] = 0.05 * torch.randn(shape, generator=g)
        else:
            fan_in = int(np.prod(shape[1:]))
            sd[name] = torch.randn(shape, generator=g) / np.sqrt(fan_in)
    return sd


def _bn(x: torch.Tensor, sd: SD, p: str) -> torch.Tensor:
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"], False, 0.0, 1e-5)


def bcvae_encode_mean(sd: SD, coeffs: torch.Tensor) -> torch.Tensor:
    """``BCVAE.encode(coeffs).mean`` in eval mode (``vae.py:68-89``): coeffs (N, 120, 32) -> (N, 64)."""
    x = coeffs.transpose(1, 2)
    x = F.leaky_relu(_bn(F.conv1d(x, sd["encoder.conv_layers.0.weight"], sd["encoder.conv_layers.0.bias"]), sd, "encoder.conv_layers.1"), 0.2)
    x = F.leaky_relu(_bn(F.conv1d(x, sd["encoder.conv_layers.3.weight"], sd["encoder.conv_layers.3.bias"]), sd, "encoder.conv_layers.4"), 0.2)
    x = F.leaky_relu(_bn(F.conv1d(x, sd["encoder.conv_layers.6.weight"], sd["encoder.conv_layers.6.bias"], stride=2), sd, "encoder.conv_layers.7"), 0.2)
    x = F.conv1d(x, sd["encoder.conv_layers.9.weight"], sd["encoder.conv_layers.9.bias"]).flatten(1)
    x = F.leaky_relu(_bn(F.linear(x, sd["encoder.fc_layers.0.weight"], sd["encoder.fc_layers.0.bias"]), sd, "encoder.fc_layers.1"), 0.01)
    x = F.leaky_relu(_bn(F.linear(x, sd["encoder.fc_layers.3.weight"], sd["encoder.fc_layers.3.bias"]), sd, "encoder.fc_layers.4"), 0.01)
    x = F.linear(x, sd["encoder.fc_layers.6.weight"], sd["encoder.fc_layers.6.bias"])
    return F.linear(x, sd["encoder.fc_mu.weight"], sd["encoder.fc_mu.bias"])


def window_latents(sd: SD, coeffs: torch.Tensor, step: int) -> torch.Tensor:
    """``generate_latents_info`` (``script/test_evaluate.py:53-106``, padding 0): (B, T, 32) -> (B * num_windows, 64), window-major per clip."""
    B, T, _ = coeffs.shape
    nw = (T - SEQ_LEN) // step + 1
    wins = torch.stack([coeffs[:, s * step: s * step + SEQ_LEN] for s in range(nw)], dim=1)   # (B, nw, 120, 32)
    return bcvae_encode_mean(sd, wins.reshape(B * nw, SEQ_LEN, -1))


def frechet_distance(lat1: np.ndarray, lat2: np.ndarray) -> float:
    """``get_statistic`` + ``frechet_distance`` (``said/metric/frechet_distance.py:17-64`` -> pytorch_fid ``calculate_frechet_distance``)."""
    from scipy import linalg

    mu1, mu2 = np.mean(lat1, axis=0), np.mean(lat2, axis=0)
    s1, s2 = np.cov(lat1, rowvar=False), np.cov(lat2, rowvar=False)
    diff = mu1 - mu2
    covmean = linalg.sqrtm(s1.dot(s2))
    if not np.isfinite(covmean).all():
        off = np.eye(s1.shape[0]) * 1e-6
        covmean = linalg.sqrtm((s1 + off).dot(s2 + off))
    if np.iscomplexobj(covmean):
        covmean = covmean.real
    return float(diff.dot(diff) + np.trace(s1) + np.trace(s2) - 2 * np.trace(covmean))
