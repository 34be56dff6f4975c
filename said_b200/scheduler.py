"""Host side of the noise scheduler used by ``SAID.inference``.

The reference builds ``diffusers.DDIMScheduler(num_train_timesteps, beta_schedule="squaredcos_cap_v2",
prediction_type=...)`` with every other knob at its diffusers-0.19 default
(``said/model/diffusion.py:100-104``) and calls ``set_timesteps`` / ``scale_model_input`` / ``step`` /
``add_noise`` on it (``diffusion.py:361, 424-426, 441-443, 271-272, 452-454``).  diffusers is a pinned
third-party dependency (``pyproject.toml:16``, ``==0.19.*``) that is not vendored in the reference, so
this file restates the published algorithm (diffusers v0.19.3
``src/diffusers/schedulers/scheduling_ddim.py``) for the pieces that stay on the host:

* the cumulative-alpha table (float64 cosine -> float32 betas -> float32 cumprod),
* the "leading" timestep grid,
* the per-step coefficient table handed to the fused CUDA step kernel.

The per-element step arithmetic itself runs on the GPU (``csrc/diffusion_kernels.cuh``).  ``step`` /
``add_noise`` / ``get_velocity`` below are the same formulas in torch, kept for the non-hot-path API
(``SAID.add_noise``, ``SAID.pred_original_sample``; training uses them) -- ``inference()`` never calls
them.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from types import SimpleNamespace
from typing import List, Optional, Sequence, Union

import numpy as np
import torch

PRED_EPSILON, PRED_SAMPLE, PRED_V = 0, 1, 2
_PRED_CODES = {"epsilon": PRED_EPSILON, "sample": PRED_SAMPLE, "v_prediction": PRED_V}


def cosine_alphas_cumprod(num_train_timesteps: int, max_beta: float = 0.999) -> torch.Tensor:
    """``betas_for_alpha_bar`` (cosine) -> ``cumprod(1 - betas)`` in float32, as diffusers does."""

    def alpha_bar(s: float) -> float:
        return math.cos((s + 0.008) / 1.008 * math.pi / 2) ** 2

    n = num_train_timesteps
    betas = [min(1.0 - alpha_bar((i + 1) / n) / alpha_bar(i / n), max_beta) for i in range(n)]
    betas_t = torch.tensor(betas, dtype=torch.float32)
    return torch.cumprod(1.0 - betas_t, dim=0)


@dataclass
class SchedulerOutput:
    prev_sample: torch.Tensor
    pred_original_sample: Optional[torch.Tensor] = None


def noise_coefs(scheduler, t: int):
    """``(sqrt(abar[t]), sqrt(1 - abar[t]))`` evaluated like ``DDIMScheduler.add_noise`` does (float32
    tensor ``** 0.5``).  Works on any scheduler object exposing ``alphas_cumprod``."""
    a = scheduler.alphas_cumprod.detach().to("cpu", torch.float32)[int(t)]
    return float(a**0.5), float((1 - a) ** 0.5)


def ddim_step_table(scheduler, timesteps: Sequence[int], eta: float = 0.0,
                    blend_next: Optional[Sequence[Optional[int]]] = None) -> np.ndarray:
    """Per-step scalars for the fused CUDA step kernel (``csrc/diffusion_kernels.cuh``), float32, computed
    with the same float32 operation order as ``DDIMScheduler.step`` (0-d float32 tensors, ``** 0.5``).

    Row ``k`` = ``[sqrt_a, sqrt_b, sqrt_ap, dir_coef, sigma, clip, blend_sa, blend_sb]`` with
    ``a = abar[t]``, ``b = 1 - a``, ``ap = abar[t_prev]`` (``final_alpha_cumprod`` below 0),
    ``sigma = eta * sqrt((1-ap)/(1-a) * (1 - a/ap))``, ``dir_coef = sqrt(1 - ap - sigma^2)``,
    ``clip`` = ``clip_sample_range`` or -1 when clipping is off, and ``blend_*`` the ``add_noise``
    coefficients at ``blend_next[k]`` (the timestep of the following iteration, used by the editing
    blend, ``said/model/diffusion.py:446-456``; ``None`` = last iteration = un-noised init).

    ``scheduler`` may be this module's :class:`DDIMScheduler` or diffusers' (same attribute names).
    """
    n = len(timesteps)
    rows = np.zeros((n, 8), dtype=np.float32)
    if n == 0:
        return rows
    ac = scheduler.alphas_cumprod.detach().to("cpu", torch.float32)
    cfg = scheduler.config
    final = torch.as_tensor(scheduler.final_alpha_cumprod, dtype=torch.float32).to("cpu")
    clip = float(getattr(cfg, "clip_sample_range", 1.0)) if cfg.clip_sample else -1.0
    ratio = cfg.num_train_timesteps // int(scheduler.num_inference_steps)
    # vectorised over steps; every op is an IEEE float32 elementwise op, so the values equal the 0-d
    # tensor arithmetic of DDIMScheduler.step bit for bit
    ts = torch.as_tensor(np.asarray(timesteps, dtype=np.int64))
    tp = ts - ratio
    a = ac[ts]
    ap = torch.where(tp >= 0, ac[tp.clamp(min=0)], final)
    b = 1 - a
    variance = ((1 - ap) / (1 - a)) * (1 - a / ap)
    sigma = float(eta) * variance**0.5
    dir_coef = (1 - ap - sigma**2) ** 0.5
    rows[:, 0] = (a**0.5).numpy()
    rows[:, 1] = (b**0.5).numpy()
    rows[:, 2] = (ap**0.5).numpy()
    rows[:, 3] = dir_coef.numpy()
    rows[:, 4] = sigma.numpy()
    rows[:, 5] = clip
    rows[:, 6], rows[:, 7] = 1.0, 0.0
    if blend_next is not None:
        idx = [k for k in range(n) if blend_next[k] is not None]
        if idx:
            an = ac[torch.as_tensor([int(blend_next[k]) for k in idx], dtype=torch.long)]
            rows[idx, 6] = (an**0.5).numpy()
            rows[idx, 7] = ((1 - an) ** 0.5).numpy()
    return rows


class SchedulerMixin:
    """Type tag mirroring ``diffusers.SchedulerMixin`` (the reference only uses it in annotations)."""


class DDIMScheduler(SchedulerMixin):
    """diffusers-0.19 ``DDIMScheduler`` restricted to the configuration SAiD instantiates."""

    order = 1

    def __init__(
        self,
        num_train_timesteps: int = 1000,
        beta_start: float = 0.0001,
        beta_end: float = 0.02,
        beta_schedule: str = "linear",
        trained_betas=None,
        clip_sample: bool = True,
        set_alpha_to_one: bool = True,
        steps_offset: int = 0,
        prediction_type: str = "epsilon",
        thresholding: bool = False,
        dynamic_thresholding_ratio: float = 0.995,
        clip_sample_range: float = 1.0,
        sample_max_value: float = 1.0,
        timestep_spacing: str = "leading",
        rescale_betas_zero_snr: bool = False,
    ):
        if trained_betas is not None:
            betas = torch.tensor(trained_betas, dtype=torch.float32)
            self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        elif beta_schedule == "squaredcos_cap_v2":
            self.alphas_cumprod = cosine_alphas_cumprod(num_train_timesteps)
        elif beta_schedule == "linear":
            betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
            self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        elif beta_schedule == "scaled_linear":
            betas = torch.linspace(beta_start**0.5, beta_end**0.5, num_train_timesteps, dtype=torch.float32) ** 2
            self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        else:
            raise NotImplementedError(f"{beta_schedule} is not implemented for {self.__class__}")
        if thresholding or rescale_betas_zero_snr:
            raise NotImplementedError("thresholding / rescale_betas_zero_snr are not used by SAiD")
        if prediction_type not in _PRED_CODES:
            raise ValueError(f"prediction_type given as {prediction_type} must be one of {list(_PRED_CODES)}")
        if timestep_spacing not in ("leading", "trailing", "linspace"):
            raise ValueError(f"{timestep_spacing} is not supported")
        self.config = SimpleNamespace(
            num_train_timesteps=num_train_timesteps,
            beta_schedule=beta_schedule,
            clip_sample=clip_sample,
            set_alpha_to_one=set_alpha_to_one,
            steps_offset=steps_offset,
            prediction_type=prediction_type,
            clip_sample_range=clip_sample_range,
            timestep_spacing=timestep_spacing,
        )
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.init_noise_sigma = 1.0
        self.num_inference_steps: Optional[int] = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))

    # ------------------------------------------------------------------ grid
    def set_timesteps(self, num_inference_steps: int, device: Union[str, torch.device, None] = None) -> None:
        n_train = self.config.num_train_timesteps
        if num_inference_steps > n_train:
            raise ValueError(
                f"`num_inference_steps`: {num_inference_steps} cannot be larger than "
                f"`self.config.train_timesteps`: {n_train}"
            )
        self.num_inference_steps = num_inference_steps
        spacing = self.config.timestep_spacing
        if spacing == "leading":
            ratio = n_train // num_inference_steps
            ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64)
            ts += self.config.steps_offset
        elif spacing == "trailing":
            ratio = n_train / num_inference_steps
            ts = np.round(np.arange(n_train, 0, -ratio)).astype(np.int64) - 1
        else:  # linspace
            ts = np.linspace(0, n_train - 1, num_inference_steps).round()[::-1].copy().astype(np.int64)
        self._timesteps_host = ts          # host copy: the engine's step table is built without a device sync
        self.timesteps = torch.from_numpy(ts).to(device)

    def scale_model_input(self, sample: torch.Tensor, timestep=None) -> torch.Tensor:
        return sample

    def prev_timestep(self, t: int) -> int:
        return int(t) - self.config.num_train_timesteps // int(self.num_inference_steps)

    # ------------------------------------------------- coefficient table for the CUDA step kernel
    def step_table(self, timesteps: List[int], eta: float = 0.0) -> np.ndarray:
        """See :func:`ddim_step_table` (blend columns left at the identity)."""
        return ddim_step_table(self, timesteps, eta)

    def noise_coefs(self, t: int):
        """``(sqrt(abar[t]), sqrt(1-abar[t]))`` as float32 scalars (``add_noise`` operation order)."""
        return noise_coefs(self, t)

    @property
    def prediction_code(self) -> int:
        return _PRED_CODES[self.config.prediction_type]

    # ----------------------------------------------- torch formulas (non-hot-path API parity)
    def step(
        self,
        model_output: torch.Tensor,
        timestep: int,
        sample: torch.Tensor,
        eta: float = 0.0,
        use_clipped_model_output: bool = False,
        generator=None,
        variance_noise: Optional[torch.Tensor] = None,
        return_dict: bool = True,
    ):
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' first")
        t = int(timestep)
        tp = self.prev_timestep(t)
        ac = self.alphas_cumprod
        a = ac[t]
        ap = ac[tp] if tp >= 0 else self.final_alpha_cumprod
        b = 1 - a
        p = self.config.prediction_type
        if p == "epsilon":
            x0 = (sample - b**0.5 * model_output) / a**0.5
            eps = model_output
        elif p == "sample":
            x0 = model_output
            eps = (sample - a**0.5 * x0) / b**0.5
        else:
            x0 = (a**0.5) * sample - (b**0.5) * model_output
            eps = (a**0.5) * model_output + (b**0.5) * sample
        if self.config.clip_sample:
            x0 = x0.clamp(-self.config.clip_sample_range, self.config.clip_sample_range)
        variance = ((1 - ap) / (1 - a)) * (1 - a / ap)
        std = eta * variance**0.5
        if use_clipped_model_output:
            eps = (sample - a**0.5 * x0) / b**0.5
        direction = (1 - ap - std**2) ** 0.5 * eps
        prev = ap**0.5 * x0 + direction
        if eta > 0:
            if variance_noise is None:
                variance_noise = torch.randn(
                    model_output.shape, generator=generator, device=model_output.device, dtype=model_output.dtype
                )
            prev = prev + std * variance_noise
        if not return_dict:
            return (prev,)
        return SchedulerOutput(prev_sample=prev, pred_original_sample=x0)

    def _coefs(self, like: torch.Tensor, timesteps: torch.Tensor):
        ac = self.alphas_cumprod.to(device=like.device, dtype=like.dtype)
        timesteps = timesteps.to(like.device)
        sa = ac[timesteps] ** 0.5
        sb = (1 - ac[timesteps]) ** 0.5
        sa = sa.flatten()
        sb = sb.flatten()
        while sa.dim() < like.dim():
            sa = sa.unsqueeze(-1)
            sb = sb.unsqueeze(-1)
        return sa, sb

    def add_noise(self, original_samples: torch.Tensor, noise: torch.Tensor, timesteps: torch.Tensor) -> torch.Tensor:
        sa, sb = self._coefs(original_samples, timesteps)
        return sa * original_samples + sb * noise

    def get_velocity(self, sample: torch.Tensor, noise: torch.Tensor, timesteps: torch.Tensor) -> torch.Tensor:
        sa, sb = self._coefs(sample, timesteps)
        return sa * noise - sb * sample

    def __len__(self) -> int:
        return self.config.num_train_timesteps


# ======================================================================================================
# DDPMScheduler (ancestral sampling).  The reference's constructor takes ``noise_scheduler: Type[SchedulerMixin]``
# (``said/model/diffusion.py:55``) and calls only ``set_timesteps / scale_model_input / step / add_noise /
# get_velocity`` on it; ``DDPMScheduler.step`` has no ``eta`` argument, so the ``inspect.signature`` test of
# ``diffusion.py:404-405`` leaves ``extra_step_kwargs`` empty.  Restated from diffusers v0.19.3
# ``src/diffusers/schedulers/scheduling_ddpm.py`` (same caveat as above: the package is absent, parity unpinned).
# ======================================================================================================
def ddpm_step_table(scheduler, timesteps: Sequence[int], blend_next: Optional[Sequence[Optional[int]]] = None) -> np.ndarray:
    """Per-step scalars of ``DDPMScheduler.step`` for the fused CUDA step kernel (scheduler code 1), float32 with the
    operation order of diffusers' 0-d tensor arithmetic.

    Row ``k`` = ``[sqrt_a, sqrt_b, c0, c1, std, clip, blend_sa, blend_sb]`` with ``a = abar[t]``, ``ap = abar[t_prev]``
    (1 below 0), ``cur_a = a / ap``, ``cur_b = 1 - cur_a``, ``c0 = sqrt(ap) * cur_b / (1 - a)`` (pred_original_sample_coeff),
    ``c1 = sqrt(cur_a) * (1 - ap) / (1 - a)`` (current_sample_coeff), ``std = sqrt(clamp((1-ap)/(1-a) * cur_b, 1e-20))`` for
    ``t > 0`` and 0 at ``t = 0`` (no variance noise is drawn there)."""
    n = len(timesteps)
    rows = np.zeros((n, 8), dtype=np.float32)
    if n == 0:
        return rows
    if scheduler.config.variance_type != "fixed_small":
        raise NotImplementedError("DDPM variance_type other than 'fixed_small' is not supported on the CUDA path")
    ac = scheduler.alphas_cumprod.detach().to("cpu", torch.float32)
    cfg = scheduler.config
    one = torch.tensor(1.0, dtype=torch.float32)
    clip = float(getattr(cfg, "clip_sample_range", 1.0)) if cfg.clip_sample else -1.0
    ratio = cfg.num_train_timesteps // int(scheduler.num_inference_steps)
    ts = torch.as_tensor(np.asarray(timesteps, dtype=np.int64))
    tp = ts - ratio
    a = ac[ts]
    ap = torch.where(tp >= 0, ac[tp.clamp(min=0)], one)
    b = 1 - a
    bp = 1 - ap
    cur_a = a / ap
    cur_b = 1 - cur_a
    c0 = (ap**0.5 * cur_b) / b
    c1 = cur_a**0.5 * bp / b
    var = torch.clamp((1 - ap) / (1 - a) * cur_b, min=1e-20)
    std = torch.where(ts > 0, var**0.5, torch.zeros_like(var))
    rows[:, 0] = (a**0.5).numpy()
    rows[:, 1] = (b**0.5).numpy()
    rows[:, 2] = c0.numpy()
    rows[:, 3] = c1.numpy()
    rows[:, 4] = std.numpy()
    rows[:, 5] = clip
    rows[:, 6], rows[:, 7] = 1.0, 0.0
    if blend_next is not None:
        idx = [k for k in range(n) if blend_next[k] is not None]
        if idx:
            an = ac[torch.as_tensor([int(blend_next[k]) for k in idx], dtype=torch.long)]
            rows[idx, 6] = (an**0.5).numpy()
            rows[idx, 7] = ((1 - an) ** 0.5).numpy()
    return rows


class DDPMScheduler(DDIMScheduler):
    """diffusers-0.19 ``DDPMScheduler`` restricted to what SAiD can instantiate: ``squaredcos_cap_v2`` / linear betas,
    ``variance_type="fixed_small"``, ``clip_sample=True``, leading timestep grid.  Shares the alpha table, the grid,
    ``add_noise`` and ``get_velocity`` with :class:`DDIMScheduler`; ``step`` is the ancestral update."""

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.0001, beta_end: float = 0.02,
                 beta_schedule: str = "linear", trained_betas=None, variance_type: str = "fixed_small",
                 clip_sample: bool = True, prediction_type: str = "epsilon", thresholding: bool = False,
                 dynamic_thresholding_ratio: float = 0.995, clip_sample_range: float = 1.0, sample_max_value: float = 1.0,
                 timestep_spacing: str = "leading", steps_offset: int = 0):
        super().__init__(num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end,
                         beta_schedule=beta_schedule, trained_betas=trained_betas, clip_sample=clip_sample,
                         steps_offset=steps_offset, prediction_type=prediction_type, thresholding=thresholding,
                         clip_sample_range=clip_sample_range, timestep_spacing=timestep_spacing)
        if variance_type not in ("fixed_small", "fixed_small_log", "fixed_large", "fixed_large_log"):
            raise NotImplementedError(f"variance_type {variance_type} (learned variances) is not used by SAiD")
        self.config.variance_type = variance_type
        self.one = torch.tensor(1.0)

    def step_table(self, timesteps: List[int], eta: float = 0.0) -> np.ndarray:
        return ddpm_step_table(self, timesteps)

    def _get_variance(self, t: int) -> torch.Tensor:
        tp = self.prev_timestep(t)
        a = self.alphas_cumprod[t]
        ap = self.alphas_cumprod[tp] if tp >= 0 else self.one
        cur_b = 1 - a / ap
        variance = (1 - ap) / (1 - a) * cur_b
        variance = torch.clamp(variance, min=1e-20)
        vt = self.config.variance_type
        if vt == "fixed_small_log":
            variance = torch.exp(0.5 * torch.log(variance))
        elif vt == "fixed_large":
            variance = cur_b
        elif vt == "fixed_large_log":
            variance = torch.log(cur_b)
        return variance

    def step(self, model_output: torch.Tensor, timestep: int, sample: torch.Tensor, generator=None,  # noqa: D102
             variance_noise: Optional[torch.Tensor] = None, return_dict: bool = True):
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' first")
        t = int(timestep)
        tp = self.prev_timestep(t)
        a = self.alphas_cumprod[t]
        ap = self.alphas_cumprod[tp] if tp >= 0 else self.one
        b = 1 - a
        bp = 1 - ap
        cur_a = a / ap
        cur_b = 1 - cur_a
        p = self.config.prediction_type
        if p == "epsilon":
            x0 = (sample - b**0.5 * model_output) / a**0.5
        elif p == "sample":
            x0 = model_output
        else:
            x0 = (a**0.5) * sample - (b**0.5) * model_output
        if self.config.clip_sample:
            x0 = x0.clamp(-self.config.clip_sample_range, self.config.clip_sample_range)
        c0 = (ap**0.5 * cur_b) / b
        c1 = cur_a**0.5 * bp / b
        prev = c0 * x0 + c1 * sample
        if t > 0:
            if variance_noise is None:
                variance_noise = torch.randn(model_output.shape, generator=generator, device=model_output.device,
                                             dtype=model_output.dtype)
            if self.config.variance_type == "fixed_small_log":
                variance = self._get_variance(t) * variance_noise
            else:
                variance = (self._get_variance(t) ** 0.5) * variance_noise
            prev = prev + variance
        if not return_dict:
            return (prev,)
        return SchedulerOutput(prev_sample=prev, pred_original_sample=x0)
