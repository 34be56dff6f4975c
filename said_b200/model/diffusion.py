"""``SAID`` / ``SAID_UNet1D`` -- the reference's model API (``said/model/diffusion.py``) over the sm_100a engine.

Same constructor, attributes, methods, argument meaning and return types as the reference classes
(``said/model/diffusion.py:46-527``), so ``script/inference.py`` and ``script/test_inference.py`` run
unchanged against this package (see ``compat/`` and INTEGRATION.md).  What differs is underneath:

* the module tree holds parameters only (reference names and shapes, so ``load_state_dict`` /
  ``state_dict`` / ``.to`` / ``.eval`` behave the same); the arithmetic of ``inference()``, ``forward()``
  and ``get_audio_embedding()`` runs in ``libsaid_sm100.so`` through the C ABI in
  ``include/said_b200.h``;
* ``inference()`` draws its random numbers exactly like the reference (``torch.randn`` on the input's
  device, same shapes and order: ``diffusion.py:364``, ``:270``, and one draw per step when ``eta > 0``),
  then hands the whole loop to the engine: no per-step host work, no host synchronisation;
* there is no CPU path: inputs must live on a CUDA (sm_100) device.
"""
from __future__ import annotations

import inspect
from abc import ABC
from dataclasses import dataclass
from typing import Dict, List, Optional, Type, Union

import numpy as np
import torch
from torch import nn

from .. import scheduler as sched
from .._lib import Engine
from ..scheduler import DDIMScheduler, SchedulerMixin
from .params import (
    WEIGHT_NORM_ALIASES,
    ParamTree,
    Wav2Vec2Dims,
    audio_encoder_spec,
    denoiser_spec,
    zero_init_names,
)


@dataclass
class SAIDInferenceOutput:
    """Dataclass for the inference output (reference ``diffusion.py:23-32``)"""

    result: torch.FloatTensor
    intermediates: List[torch.FloatTensor]


@dataclass
class SAIDNoiseAdditionOutput:
    """Dataclass for the noise addition output (reference ``diffusion.py:35-43``)"""

    noisy_sample: torch.FloatTensor
    noise: torch.FloatTensor
    velocity: torch.FloatTensor


class AudioProcessor:
    """Offline stand-in for ``Wav2Vec2Processor.from_pretrained("facebook/wav2vec2-base-960h")``
    (reference ``diffusion.py:90-95``; needs the network).  Implements what ``SAID.process_audio`` uses
    of it: per-utterance zero-mean / unit-variance normalisation, ``(x - mean) / sqrt(var + 1e-7)`` in
    float32 (HF ``Wav2Vec2FeatureExtractor.zero_mean_unit_var_norm``), ``sampling_rate = 16000``."""

    class _FE:
        sampling_rate = 16000
        do_normalize = True
        padding_value = 0.0

    def __init__(self):
        self.feature_extractor = self._FE()

    def __call__(self, raw_speech, sampling_rate: Optional[int] = None, return_tensors: Optional[str] = "pt", **_):
        if sampling_rate is not None and sampling_rate != self.feature_extractor.sampling_rate:
            raise ValueError(
                f"The model was trained with sampling_rate={self.feature_extractor.sampling_rate}, got {sampling_rate}"
            )
        if isinstance(raw_speech, torch.Tensor):
            raw_speech = raw_speech.detach().cpu().numpy()
        batched = isinstance(raw_speech, (list, tuple)) and len(raw_speech) > 0 and isinstance(
            raw_speech[0], (np.ndarray, list, tuple, torch.Tensor)
        ) or (isinstance(raw_speech, np.ndarray) and raw_speech.ndim > 1)
        rows = list(raw_speech) if batched else [raw_speech]
        out = []
        for r in rows:
            if isinstance(r, torch.Tensor):
                r = r.detach().cpu().numpy()
            x = np.asarray(r, dtype=np.float32)
            out.append(((x - x.mean()) / np.sqrt(x.var() + 1e-7)).astype(np.float32))
        if len({o.shape for o in out}) != 1:
            raise ValueError("all waveforms of a batch must have the same length (no padding is applied)")
        arr = np.stack(out, 0)
        return {"input_values": torch.from_numpy(arr) if return_tensors == "pt" else arr}


class UNet1DConditionModel(nn.Module):
    """Parameter container with the reference's names (``said/model/unet_1d_condition.py:12-49``)."""

    def __init__(self, in_channels: int, out_channels: int, cross_attention_dim: int):
        super().__init__()
        if in_channels != out_channels:
            raise NotImplementedError("the SAiD denoiser has in_channels == out_channels")
        self.in_channels = in_channels
        self.cross_attention_dim = cross_attention_dim
        self.model = ParamTree(denoiser_spec(in_channels, cross_attention_dim))

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("call SAID.forward(); the denoiser arithmetic runs in the sm_100a engine")


def _init_like_reference(module: nn.Module, zero_names: set, prefix: str) -> None:
    """Fresh (untrained) parameters: small random weights, unit norms, and the reference's zero-initialised
    output layers (``zero_module``, ``openaimodel.py:182-185, 668``; ``attention.py:221``)."""
    with torch.no_grad():
        for name, p in module.named_parameters():
            full = prefix + name
            leaf = name.rsplit(".", 1)[-1]
            if full in zero_names:
                p.zero_()
            elif "norm" in name or name.startswith("out.0.") or ".in_layers.0." in name or ".out_layers.0." in name:
                p.fill_(1.0 if leaf == "weight" else 0.0)
            elif leaf == "weight_g":
                p.fill_(1.0)
            elif leaf == "bias":
                p.zero_()
            else:
                p.normal_(0.0, 0.02)


class SAID(ABC, nn.Module):
    """Abstract class of SAiD models (reference ``diffusion.py:46-472``)"""

    denoiser: nn.Module

    def __init__(
        self,
        audio_config=None,
        audio_processor=None,
        noise_scheduler: Type[SchedulerMixin] = DDIMScheduler,
        in_channels: int = 32,
        feature_dim: int = -1,
        diffusion_steps: int = 1000,
        latent_scale: float = 1,
        prediction_type: str = "epsilon",
    ):
        super().__init__()
        # Audio-related
        self._audio_dims = Wav2Vec2Dims(audio_config)
        # the reference stores Wav2Vec2Config() when none is given (diffusion.py:83-86): keep `model.audio_config.hidden_size` etc. readable
        self.audio_config = audio_config if audio_config is not None else self._audio_dims.as_config()
        self._audio_dims.check_supported()
        self.audio_encoder = ParamTree(audio_encoder_spec(self._audio_dims))
        self.audio_processor = audio_processor if audio_processor is not None else AudioProcessor()
        self.sampling_rate = self.audio_processor.feature_extractor.sampling_rate

        self.latent_scale = latent_scale

        # Noise scheduler
        self.noise_scheduler = noise_scheduler(
            num_train_timesteps=diffusion_steps,
            beta_schedule="squaredcos_cap_v2",
            prediction_type=prediction_type,
        )

        # Feature embedding
        self.feature_dim = feature_dim
        hidden = self._audio_dims.output_hidden
        if self.feature_dim > 0:
            self.audio_proj_layer = nn.Linear(hidden, self.feature_dim)
            self.null_cond_emb = nn.Parameter(torch.randn(1, 1, self.feature_dim))
        else:
            self.null_cond_emb = nn.Parameter(torch.randn(1, 1, hidden))

        self._engines: Dict[int, Engine] = {}
        self._engine_keys: Dict[int, tuple] = {}
        self.use_cuda_graph = True
        self.dedup_audio = True           # a batch that repeats one clip (script/test_inference.py:167) is encoded once
        self.eta_chunk_bytes = 256 << 20  # resident per-step variance noise (eta > 0 / DDPMScheduler): the loop runs in chunks of this much
        # contraction precision of the denoiser GEMMs: "tf32x3" (tcgen05, 3xTF32 split: fp32-level accuracy),
        # "fp16x3" (tcgen05 over fp16 hi/lo operand pairs loaded by TMA: same accuracy class at twice the tensor-core rate),
        # "tf32" (tcgen05, single pass) or "fp32" (FFMA); GEMMs below tc_min_rows rows stay on the FFMA kernel
        self.precision = "fp16x3"
        # the audio encoder's dense contractions: "fp16x3" (tensor cores, split-K for the long ones: 6e-5 from the reference, 4x faster)
        # at batch scale, the IEEE fp32 kernels below tc_min_rows frames (single short clips) or when set to "fp32"
        self.encoder_precision = "fp16x3"
        self.tc_min_rows = 0

    # ------------------------------------------------------------------ state dict compatibility
    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        """Accepts checkpoints written with either weight-norm spelling of the positional conv
        (transformers 4.30.2 ``weight_g/weight_v`` or the torch parametrisation names)."""
        remapped = {}
        for k, v in state_dict.items():
            for new, old in WEIGHT_NORM_ALIASES.items():
                if k == "audio_encoder." + new:
                    k = "audio_encoder." + old
            remapped[k] = v
        return super().load_state_dict(remapped, strict=strict, **kw)

    # ------------------------------------------------------------------ engine plumbing
    def _engine(self, device: torch.device) -> Engine:
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError(
                f"said_b200 runs on CUDA sm_100a devices only; got a tensor on '{device}'. "
                "Move the model and its inputs to a B200 (`.to('cuda:0')`)."
            )
        idx = device.index if device.index is not None else torch.cuda.current_device()
        eng = self._engines.get(idx)
        if eng is None:
            eng = Engine(torch.device("cuda", idx))
            self._engines[idx] = eng
        # re-upload whenever a parameter was replaced or modified in place
        key = tuple((p.data_ptr(), p._version) for p in self.parameters())
        if self._engine_keys.get(idx) != key:
            tensors = {k: v for k, v in self.state_dict().items()}
            half = 96
            tensors["time_freqs"] = torch.exp(
                -np.log(10000) * torch.arange(start=0, end=half, dtype=torch.float32) / half
            )  # ldm/util.py:75-78, evaluated with the same torch ops as the reference
            # the one encoder switch that does not show in the state-dict names
            tensors["audio_encoder.config.do_stable_layer_norm"] = torch.tensor([1.0 if self._audio_dims.stable_layer_norm else 0.0])
            eng.load_weights(tensors)
            self._engine_keys[idx] = key
        eng.set_precision(self.precision, self.tc_min_rows, self.encoder_precision)
        return eng

    # ------------------------------------------------------------------ reference API
    def forward(
        self,
        noisy_samples: torch.FloatTensor,
        timesteps: torch.LongTensor,
        audio_embedding: torch.FloatTensor,
    ) -> torch.FloatTensor:
        """Return the predicted noise in the noisy samples (reference ``diffusion.py:127-155``)

        noisy_samples (B, T, in_channels); timesteps (B,), (1,) or 0-d; audio_embedding (B, T, dim).
        """
        eng = self._engine(noisy_samples.device)
        timesteps = torch.as_tensor(timesteps)
        return eng.denoiser_forward(noisy_samples.float(), timesteps, audio_embedding.float())

    def pred_original_sample(self, noisy_samples, noise, timesteps):
        """Predict x_0 from the noisy samples and the noise (reference ``diffusion.py:157-186``)"""
        alpha_prod_t = self.noise_scheduler.alphas_cumprod[timesteps.to(self.noise_scheduler.alphas_cumprod.device)]
        alpha_prod_t = alpha_prod_t.to(noisy_samples.device).view(-1, 1, 1)
        beta_prod_t = 1 - alpha_prod_t
        return (noisy_samples - beta_prod_t**0.5 * noise) / alpha_prod_t**0.5

    def process_audio(self, waveform: Union[np.ndarray, torch.Tensor, List[np.ndarray]]) -> torch.FloatTensor:
        """Process the waveform to fit the audio encoder (reference ``diffusion.py:188-207``):
        returns (Batch_size, T_a) float32 on the CPU."""
        out = self.audio_processor(waveform, sampling_rate=self.sampling_rate, return_tensors="pt")["input_values"]
        return out

    def process_audio_device(self, waveform: torch.Tensor) -> torch.FloatTensor:
        """``process_audio`` for equal-length raw clips that already live on the GPU: (B, T_a) -> (B, T_a) on the same
        device, per-utterance zero mean / unit variance in one kernel (no host round trip; the reference's
        ``process_audio`` normalises in numpy and returns a CPU tensor, ``diffusion.py:188-207``)."""
        if waveform.dim() == 1:
            waveform = waveform[None]
        return self._engine(waveform.device).normalize_audio(waveform.float())

    def _conv_out_frames(self, n: int) -> int:
        for k, s in zip(self._audio_dims.conv_kernel, self._audio_dims.conv_stride):
            n = (n - k) // s + 1
        return n

    def get_audio_embedding(self, waveform: torch.FloatTensor, num_frames: Optional[int]) -> torch.FloatTensor:
        """Audio embedding of the waveform (reference ``diffusion.py:209-230``):
        (B, T_a) -> (B, num_frames, embed_size); ``num_frames=None`` keeps the encoder's own frame rate."""
        eng = self._engine(waveform.device)
        if num_frames is None:
            num_frames = self._conv_out_frames(waveform.shape[1])
        return eng.encode_audio(waveform.float(), int(num_frames))

    def get_random_timesteps(self, batch_size: int) -> torch.LongTensor:
        """Reference ``diffusion.py:232-251``"""
        return torch.randint(0, self.noise_scheduler.config.num_train_timesteps, (batch_size,), dtype=torch.long)

    def add_noise(self, sample: torch.FloatTensor, timestep: torch.LongTensor) -> SAIDNoiseAdditionOutput:
        """Reference ``diffusion.py:253-276``"""
        noise = torch.randn(sample.shape, device=sample.device)
        noisy_sample = self.noise_scheduler.add_noise(sample, noise, timestep)
        velocity = self.noise_scheduler.get_velocity(sample, noise, timestep)
        return SAIDNoiseAdditionOutput(noisy_sample=noisy_sample, noise=noise, velocity=velocity)

    def encode_samples(self, samples: torch.FloatTensor) -> torch.FloatTensor:
        """Reference ``diffusion.py:278-291``"""
        return samples.clone()

    def decode_latent(self, latent: torch.FloatTensor) -> torch.FloatTensor:
        """Reference ``diffusion.py:293-306``"""
        return latent.clone()

    def inference(
        self,
        waveform_processed: torch.FloatTensor,
        init_samples: Optional[torch.FloatTensor] = None,
        mask: Optional[torch.FloatTensor] = None,
        num_inference_steps: int = 100,
        strength: float = 1.0,
        guidance_scale: float = 2.5,
        guidance_rescale: float = 0.0,
        eta: float = 0.0,
        fps: int = 60,
        save_intermediate: bool = False,
        show_process: bool = False,
    ) -> SAIDInferenceOutput:
        """Inference pipeline -- same contract as the reference (``diffusion.py:308-472``).

        waveform_processed (B, T_a) processed mono waveform; init_samples / mask (B, T, in_channels) for the
        editing mode; returns ``SAIDInferenceOutput(result (B, T, in_channels) in [0, 1], intermediates)``.
        """
        batch_size = waveform_processed.shape[0]
        waveform_len = waveform_processed.shape[1]
        in_channels = self.denoiser.in_channels
        device = waveform_processed.device
        window_size = int(waveform_len / self.sampling_rate * fps)

        # random draws: same calls, shapes, device and order as the reference
        noise = None
        if init_samples is None:
            noise = torch.randn(batch_size, window_size, in_channels, device=device)      # diffusion.py:364
        else:
            noise = torch.randn(init_samples.shape, device=init_samples.device)           # diffusion.py:270
        n_loop = self._loop_length(num_inference_steps, strength)
        # the latents keep init_samples' own length in the editing mode (reference diffusion.py:366); only the audio features
        # are resampled to window_size frames
        n_frames = window_size if init_samples is None else int(init_samples.shape[1])
        # Per-step variance noise (eta > 0: DDIMScheduler.step draws randn(model_output.shape) once per iteration; DDPMScheduler.step in
        # every iteration whose timestep is > 0).  It is drawn lazily, in loop order, by the chunked loop of _run -- the same
        # torch.randn calls as the reference, but only a bounded number of steps' worth is ever resident (the reference holds one
        # step's noise at a time; n_loop x B x T x C up front would be 2.5 GB at batch 64 x 1000 steps).
        eta_noise = None
        shape = (batch_size, n_frames, in_channels)
        if eta > 0 and n_loop > 0 and self._scheduler_has_eta():
            def eta_noise(i: int, out: torch.Tensor) -> None:
                torch.randn(shape, device=device, out=out)
        elif self._scheduler_is_ddpm() and n_loop > 0:
            self.noise_scheduler.set_timesteps(num_inference_steps, device=device)
            ddpm_ts = [int(t) for t in self.noise_scheduler.timesteps.detach().cpu().numpy()][num_inference_steps - n_loop:]

            def eta_noise(i: int, out: torch.Tensor) -> None:
                if ddpm_ts[i] > 0:
                    torch.randn(shape, device=device, out=out)
                else:
                    out.zero_()
        return self._run(
            waveform_processed, noise, init_samples, mask, num_inference_steps, strength, guidance_scale,
            guidance_rescale, eta, window_size, save_intermediate, show_process, eta_noise,
        )

    # ------------------------------------------------------------------ internals
    @staticmethod
    def _loop_length(num_inference_steps: int, strength: float) -> int:
        return min(int(num_inference_steps * strength), num_inference_steps)

    def _scheduler_has_eta(self) -> bool:
        return "eta" in set(inspect.signature(self.noise_scheduler.step).parameters.keys())

    def _scheduler_is_ddpm(self) -> bool:
        """``DDPMScheduler`` (this package's or diffusers'): ancestral step, scheduler code 1 of the step kernel."""
        return type(self.noise_scheduler).__name__ == "DDPMScheduler"

    def _run(
        self,
        waveform_processed: torch.Tensor,
        noise: torch.Tensor,
        init_samples: Optional[torch.Tensor],
        mask: Optional[torch.Tensor],
        num_inference_steps: int,
        strength: float,
        guidance_scale: float,
        guidance_rescale: float,
        eta: float,
        window_size: int,
        save_intermediate: bool,
        show_process: bool,
        eta_noise: Optional[torch.Tensor],
        return_latents: bool = False,
    ) -> SAIDInferenceOutput:
        """The loop with every random tensor supplied by the caller (``inference`` draws them like the
        reference; ``said_b200.parallel`` draws them once for the whole batch and shards them)."""
        device = waveform_processed.device
        eng = self._engine(device)
        ns = self.noise_scheduler
        is_ddpm = self._scheduler_is_ddpm()
        if not hasattr(ns, "alphas_cumprod") or not (is_ddpm or hasattr(ns, "final_alpha_cumprod")):
            raise NotImplementedError(
                f"{type(ns).__name__}: the fused step kernel implements DDIMScheduler.step (what script/inference.py and "
                "script/test_inference.py construct) and DDPMScheduler.step; other schedulers are not implemented"
            )
        batch_size, in_channels = waveform_processed.shape[0], self.denoiser.in_channels
        do_cfg = guidance_scale > 1.0
        ns.set_timesteps(num_inference_steps, device=device)
        timesteps = getattr(ns, "_timesteps_host", None)
        if timesteps is None or len(timesteps) != num_inference_steps:
            timesteps = ns.timesteps.detach().cpu().numpy()
        timesteps = [int(t) for t in timesteps]

        init_timestep = self._loop_length(num_inference_steps, strength)
        t_start = num_inference_steps - init_timestep
        loop_ts = timesteps[t_start:]
        n_loop = len(loop_ts)

        editing = init_samples is not None
        src = (init_samples if editing else noise).to(device=device, dtype=torch.float32)
        # Editing: init_samples may be a few frames shorter or longer than the audio window (a previous result cut to
        # floor(len * fps / sr) frames, script/inference.py:188, against audio zero-padded by fit_audio_unet): the latents keep
        # their own length and the cross-attention uses the reference's general alignment window (attention.py:170-189).
        if src.dim() != 3 or src.shape[0] != batch_size or src.shape[2] != in_channels or (not editing and src.shape[1] != window_size):
            raise ValueError(
                f"{'init_samples' if editing else 'noise'} must be (batch={batch_size}, frames, channels={in_channels}); got {tuple(src.shape)}"
            )
        if noise.shape != src.shape:
            raise ValueError(f"noise {tuple(noise.shape)} does not match the latents {tuple(src.shape)}")
        n_frames = int(src.shape[1])
        edit_noise, edit_coefs = None, (1.0, 0.0)
        if editing:
            edit_noise = noise.to(device=device, dtype=torch.float32)
            edit_coefs = sched.noise_coefs(ns, timesteps[-init_timestep])        # diffusion.py:376-385
        use_mask = editing and mask is not None
        if use_mask:
            mask = mask.to(device=device, dtype=torch.float32)
            if mask.shape != src.shape:
                mask = mask.expand_as(src)
        blend_next = None
        if use_mask:
            blend_next = [timesteps[t_start + i + 1] if t_start + i + 1 < num_inference_steps else None for i in range(n_loop)]
        has_eta = self._scheduler_has_eta()
        if is_ddpm:
            table = sched.ddpm_step_table(ns, loop_ts, blend_next)
            if n_loop > 0 and any(t > 0 for t in loop_ts) and eta_noise is None:
                raise ValueError("DDPMScheduler needs the per-step variance noise")
        else:
            table = sched.ddim_step_table(ns, loop_ts, eta if has_eta else 0.0, blend_next)
            if eta_noise is not None and not (eta > 0 and has_eta):
                eta_noise = None
            if eta > 0 and has_eta and n_loop > 0 and eta_noise is None:
                raise ValueError("eta > 0 needs the per-step variance noise")

        # audio encoder once per clip, then the K/V hoist
        w32 = waveform_processed.to(torch.float32)
        if self.dedup_audio and batch_size > 1 and bool((w32 == w32[:1]).all()):
            # script/test_inference.py:167-168 repeats ONE clip over the batch (different noise per row): encode it once.
            # (one tiny device reduction + a host read before the loop starts)
            emb = eng.encode_audio(w32[:1].contiguous(), window_size).expand(batch_size, -1, -1).contiguous()
        else:
            emb = eng.encode_audio(w32, window_size)
        eng.prepare_context(emb, do_cfg)

        inter = None
        if save_intermediate and n_loop > 0:
            inter = torch.empty((n_loop, batch_size, n_frames, in_channels), dtype=torch.float32, device=device)
        if getattr(self, "_profile_loop", False):   # bench.py: per-kernel timing of the loop only
            eng.profile_begin()
        # The loop runs as one device program; when per-step variance noise is drawn lazily it runs in chunks of at most
        # `eta_chunk_bytes` of noise (one reused buffer), the latents carried from chunk to chunk.
        lazy = callable(eta_noise)
        chunk = n_loop
        if lazy and n_loop > 0:
            chunk = max(1, min(n_loop, int(self.eta_chunk_bytes) // (src[0].numel() * batch_size * 4)))
        noise_buf = torch.empty((chunk,) + tuple(src.shape), dtype=torch.float32, device=device) if lazy and n_loop > 0 else None
        n_chunks = max(1, -(-n_loop // max(chunk, 1)))
        latents_out = torch.empty_like(src) if (return_latents or n_chunks > 1) else None
        carry, result = src, None
        for ci in range(n_chunks):
            c0, c1 = ci * chunk, min(n_loop, (ci + 1) * chunk)
            en = eta_noise
            if lazy and n_loop > 0:
                for i in range(c0, c1):
                    eta_noise(i, noise_buf[i - c0])
                en = noise_buf[: c1 - c0]
            elif eta_noise is not None:
                en = eta_noise[c0:c1]
            result = eng.denoise(
                carry, loop_ts[c0:c1], table[c0:c1], ns_prediction_code(ns), do_cfg, guidance_scale, guidance_rescale,
                float(self.latent_scale), float(self.latent_scale) * float(ns.init_noise_sigma),
                edit_noise=edit_noise, edit_coefs=edit_coefs, mask=mask if use_mask else None,
                eta_noise=en, intermediates=None if inter is None else inter[c0:c1], latents_out=latents_out,
                use_graph=self.use_cuda_graph, scheduler=1 if is_ddpm else 0, resume=ci > 0, more=ci < n_chunks - 1,
            )
            carry = latents_out
        if not return_latents:
            latents_out = None
        eng.check_status()     # synchronises; raises if the fp16x3 path met an activation beyond fp16's range
        if show_process:
            from tqdm import tqdm

            with tqdm(total=n_loop) as bar:     # the loop is one asynchronous device program: tick once
                torch.cuda.synchronize(device)
                bar.update(n_loop)
        out = SAIDInferenceOutput(result=result, intermediates=list(inter.unbind(0)) if inter is not None else [])
        if return_latents:
            out.latents = latents_out  # type: ignore[attr-defined]
        return out


def ns_prediction_code(noise_scheduler) -> int:
    pt = noise_scheduler.config.prediction_type
    codes = {"epsilon": sched.PRED_EPSILON, "sample": sched.PRED_SAMPLE, "v_prediction": sched.PRED_V}
    if pt not in codes:
        raise ValueError(f"prediction_type given as {pt} must be one of {list(codes)}")
    return codes[pt]


class SAID_UNet1D(SAID):
    """SAiD model implemented using U-Net 1D model (reference ``diffusion.py:475-527``)"""

    def __init__(
        self,
        audio_config=None,
        audio_processor=None,
        noise_scheduler: Type[SchedulerMixin] = DDIMScheduler,
        in_channels: int = 32,
        feature_dim: int = -1,
        diffusion_steps: int = 1000,
        latent_scale: float = 1,
        prediction_type: str = "epsilon",
    ):
        super().__init__(
            audio_config=audio_config,
            audio_processor=audio_processor,
            in_channels=in_channels,
            feature_dim=feature_dim,
            diffusion_steps=diffusion_steps,
            latent_scale=latent_scale,
            prediction_type=prediction_type,
            noise_scheduler=noise_scheduler,
        )
        ctx = self.feature_dim if self.feature_dim > 0 else self._audio_dims.hidden
        self.denoiser = UNet1DConditionModel(in_channels=in_channels, out_channels=in_channels, cross_attention_dim=ctx)
        _init_like_reference(self.denoiser, set(zero_init_names()), "denoiser.")
        _init_like_reference(self.audio_encoder, set(), "audio_encoder.")
