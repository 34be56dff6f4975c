"""Distribution-level evaluation on the device.

The reference scores generated coefficient sequences by embedding sliding 120-frame windows with its BCVAE encoder
(``said/model/vae.py:26-89``; windows as ``script/test_evaluate.py:53-106``) and taking the Frechet distance between the latent
sets (``said/metric/frechet_distance.py:17-64`` -> ``pytorch_fid.calculate_frechet_distance``).  Here both steps are kernels of
``libsaid_sm100.so`` (``csrc/eval_kernels.cuh``) behind the C ABI (``said_eval_*``), so two runs of the inference path can be
compared as DISTRIBUTIONS without leaving the GPU -- the only meaningful comparison for epsilon-prediction 1000-step chains,
which are chaotic sample by sample.
"""
from __future__ import annotations

from typing import Dict

import torch

from ._lib import Engine


class DeviceEvaluator:
    """BCVAE window embedder + Frechet distance on one CUDA device."""

    def __init__(self, device, bcvae_state_dict: Dict[str, torch.Tensor]):
        self.engine = Engine(torch.device(device))
        self.engine.load_bcvae(bcvae_state_dict)

    def latents(self, coeffs: torch.Tensor, window_step: int = 30) -> torch.Tensor:
        """(B, T, 32) coefficient sequences -> (B * num_windows, 64) latent means (``BCVAE.encode(window).mean``)."""
        return self.engine.bcvae_latents(coeffs.to(self.engine.device, torch.float32), window_step)

    def frechet_distance(self, coeffs_a: torch.Tensor, coeffs_b: torch.Tensor, window_step: int = 30) -> float:
        """Frechet distance between the window-latent distributions of two sets of coefficient sequences."""
        return self.engine.frechet(self.latents(coeffs_a, window_step), self.latents(coeffs_b, window_step))["frechet_distance"]
