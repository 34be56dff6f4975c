"""CPU oracle for the SAiD inference hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs
may import this module.  The product path (``said_b200``) never does; it fails loudly when the CUDA
library is missing instead of falling back to anything in here.

What it is: a plain-PyTorch (CPU, eager) restatement of the reference's algorithm for the path
``SAID.inference()`` (``/root/reference/said/model/diffusion.py:308-472``), written as pure functions
over a reference-layout ``state_dict`` so that it runs where ``/root/reference`` does not exist (the GPU
box).  It deliberately keeps the reference's *cost structure* -- K/V re-projected from the 768-wide
context on every step, a full T x T masked cross-attention, the per-row mask construction loop -- so
that timing it is a fair stand-in for the reference's CPU path.

How it is pinned (``tests/golden/make_golden.py``, run in the build container where the reference is
mounted):
  * denoiser, audio encoder, and the whole ``inference()`` loop are checked against the reference's own
    modules (``said/model/{unet_1d_condition,wav2vec2,diffusion}.py`` imported unchanged from
    /root/reference, transformers 5.5.0 supplying ``Wav2Vec2Model``) on seeded synthetic weights, and the
    reference's outputs are committed as golden vectors under ``tests/golden/``;
  * the scheduler arithmetic is third-party: ``diffusers==0.19.*`` (``pyproject.toml:16``) is NOT
    installed here and not vendored in the reference, and the reference has no tests or golden vectors
    for it.  ``ddim_*`` / ``rescale_noise_cfg`` below restate the published v0.19.3 algorithm
    (``src/diffusers/schedulers/scheduling_ddim.py``,
    ``src/diffusers/pipelines/stable_diffusion/pipeline_stable_diffusion.py``) => PARITY UNPINNED for
    that boundary; anchors are the reference's call sites (``diffusion.py:100-104, 271-272, 361, 370,
    424-426, 436-443, 452-454``) and the known-answer values in SURVEY.md Appendix B.

All functions take ``dtype`` implicitly from the tensors they are given: pass a float64 state dict and
float64 inputs to get the fp64 oracle (the reference hard-casts to float32 in three places,
``ldm/util.py:79-82, 120-122``, ``ldm/openaimodel.py:457``; the oracle does not).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]

# --------------------------------------------------------------------------------------------------
# Scheduler (diffusers 0.19 DDIMScheduler, restated; see module docstring)
# --------------------------------------------------------------------------------------------------


def ddim_alphas_cumprod(num_train_timesteps: int = 1000) -> torch.Tensor:
    """``squaredcos_cap_v2``: betas from the float64 cosine, stored float32, cumprod in float32."""
    f = lambda s: math.cos((s + 0.008) / 1.008 * math.pi / 2) ** 2  # noqa: E731
    n = num_train_timesteps
    betas = torch.tensor([min(1 - f((i + 1) / n) / f(i / n), 0.999) for i in range(n)], dtype=torch.float32)
    return torch.cumprod(1.0 - betas, dim=0)


def ddim_timesteps(num_inference_steps: int, num_train_timesteps: int = 1000) -> np.ndarray:
    """``timestep_spacing="leading"``, ``steps_offset=0``."""
    ratio = num_train_timesteps // num_inference_steps
    return (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64)


def ddim_step(
    model_output: torch.Tensor,
    t: int,
    sample: torch.Tensor,
    alphas_cumprod: torch.Tensor,
    num_inference_steps: int,
    prediction_type: str = "epsilon",
    eta: float = 0.0,
    variance_noise: Optional[torch.Tensor] = None,
    num_train_timesteps: int = 1000,
    clip_sample: bool = True,
    clip_sample_range: float = 1.0,
) -> torch.Tensor:
    """One ``DDIMScheduler.step(...).prev_sample`` with ``use_clipped_model_output=False``."""
    t = int(t)
    t_prev = t - num_train_timesteps // num_inference_steps
    ac = alphas_cumprod.to(sample.dtype) if sample.dtype == torch.float64 else alphas_cumprod
    a = ac[t]
    a_prev = ac[t_prev] if t_prev >= 0 else torch.tensor(1.0, dtype=ac.dtype)
    b = 1 - a
    if prediction_type == "epsilon":
        x0 = (sample - b**0.5 * model_output) / a**0.5
        eps = model_output
    elif prediction_type == "sample":
        x0 = model_output
        eps = (sample - a**0.5 * x0) / b**0.5
    elif prediction_type == "v_prediction":
        x0 = (a**0.5) * sample - (b**0.5) * model_output
        eps = (a**0.5) * model_output + (b**0.5) * sample
    else:
        raise ValueError(prediction_type)
    if clip_sample:
        x0 = x0.clamp(-clip_sample_range, clip_sample_range)
    variance = ((1 - a_prev) / (1 - a)) * (1 - a / a_prev)
    std = eta * variance**0.5
    direction = (1 - a_prev - std**2) ** 0.5 * eps
    prev = a_prev**0.5 * x0 + direction
    if eta > 0:
        if variance_noise is None:
            variance_noise = torch.randn(model_output.shape, dtype=model_output.dtype)
        prev = prev + std * variance_noise
    return prev


def ddpm_step(
    model_output: torch.Tensor,
    t: int,
    sample: torch.Tensor,
    alphas_cumprod: torch.Tensor,
    num_inference_steps: int,
    prediction_type: str = "epsilon",
    variance_noise: Optional[torch.Tensor] = None,
    num_train_timesteps: int = 1000,
    clip_sample: bool = True,
    clip_sample_range: float = 1.0,
) -> torch.Tensor:
    """One ``DDPMScheduler.step(...).prev_sample`` (diffusers 0.19 ``scheduling_ddpm.py``, ``variance_type="fixed_small"``):
    the scheduler the reference gets when constructed with ``noise_scheduler=DDPMScheduler`` (``said/model/diffusion.py:55``).
    Third-party arithmetic restated from the published source: parity unpinned, like ``ddim_step``."""
    t = int(t)
    t_prev = t - num_train_timesteps // num_inference_steps
    ac = alphas_cumprod.to(sample.dtype) if sample.dtype == torch.float64 else alphas_cumprod
    a = ac[t]
    a_prev = ac[t_prev] if t_prev >= 0 else torch.tensor(1.0, dtype=ac.dtype)
    b = 1 - a
    b_prev = 1 - a_prev
    cur_a = a / a_prev
    cur_b = 1 - cur_a
    if prediction_type == "epsilon":
        x0 = (sample - b**0.5 * model_output) / a**0.5
    elif prediction_type == "sample":
        x0 = model_output
    elif prediction_type == "v_prediction":
        x0 = (a**0.5) * sample - (b**0.5) * model_output
    else:
        raise ValueError(prediction_type)
    if clip_sample:
        x0 = x0.clamp(-clip_sample_range, clip_sample_range)
    c0 = (a_prev**0.5 * cur_b) / b
    c1 = cur_a**0.5 * b_prev / b
    prev = c0 * x0 + c1 * sample
    if t > 0:
        if variance_noise is None:
            variance_noise = torch.randn(model_output.shape, dtype=model_output.dtype)
        variance = torch.clamp((1 - a_prev) / (1 - a) * cur_b, min=1e-20)
        prev = prev + (variance**0.5) * variance_noise
    return prev


def ddim_add_noise(x: torch.Tensor, noise: torch.Tensor, t, alphas_cumprod: torch.Tensor) -> torch.Tensor:
    ac = alphas_cumprod.to(dtype=x.dtype)
    t = torch.as_tensor(t, dtype=torch.long).reshape(-1)
    sa = (ac[t] ** 0.5).reshape(-1, *([1] * (x.dim() - 1))).to(x.device)
    sb = ((1 - ac[t]) ** 0.5).reshape(-1, *([1] * (x.dim() - 1))).to(x.device)
    return sa * x + sb * noise


def rescale_noise_cfg(noise_cfg: torch.Tensor, noise_pred_text: torch.Tensor, guidance_rescale: float) -> torch.Tensor:
    dims = list(range(1, noise_pred_text.ndim))
    std_text = noise_pred_text.std(dim=dims, keepdim=True)
    std_cfg = noise_cfg.std(dim=dims, keepdim=True)
    rescaled = noise_cfg * (std_text / std_cfg)
    return guidance_rescale * rescaled + (1 - guidance_rescale) * noise_cfg


# --------------------------------------------------------------------------------------------------
# Denoiser (said/model/unet_1d_condition.py + said/model/ldm/*)
# --------------------------------------------------------------------------------------------------


def timestep_embedding(timesteps: torch.Tensor, dim: int, dtype: torch.dtype) -> torch.Tensor:
    """``ldm/util.py:66-90``: ``[cos(t f) | sin(t f)]``, ``f_k = exp(-ln(1e4) k / half)``."""
    half = dim // 2
    freqs = torch.exp(-math.log(10000) * torch.arange(0, half, dtype=torch.float32) / half).to(timesteps.device)
    if dtype == torch.float64:
        freqs = torch.exp(-math.log(10000) * torch.arange(0, half, dtype=torch.float64) / half).to(timesteps.device)
        args = timesteps[:, None].to(torch.float64) * freqs[None]
    else:
        args = timesteps[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def _group_norm(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, eps: float) -> torch.Tensor:
    return F.group_norm(x, 32, w, b, eps)


def _resblock(sd: SD, p: str, x: torch.Tensor, emb: torch.Tensor) -> torch.Tensor:
    """``ResBlock._forward`` (``openaimodel.py:207-227``), no up/down, no scale-shift norm."""
    h = F.silu(_group_norm(x, sd[p + "in_layers.0.weight"], sd[p + "in_layers.0.bias"], 1e-5))
    h = F.conv1d(h, sd[p + "in_layers.2.weight"], sd[p + "in_layers.2.bias"], padding=1)
    e = F.linear(F.silu(emb), sd[p + "emb_layers.1.weight"], sd[p + "emb_layers.1.bias"])
    h = h + e[:, :, None]
    h = F.silu(_group_norm(h, sd[p + "out_layers.0.weight"], sd[p + "out_layers.0.bias"], 1e-5))
    h = F.conv1d(h, sd[p + "out_layers.3.weight"], sd[p + "out_layers.3.bias"], padding=1)
    if (p + "skip_connection.weight") in sd:
        x = F.conv1d(x, sd[p + "skip_connection.weight"], sd[p + "skip_connection.bias"])
    return x + h


def alignment_mask(batch: int, x_len: int, c_len: int, pad: int = 1, device=None) -> torch.Tensor:
    """``BasicTransformerBlock._forward`` alignment bias (``attention.py:170-189``): True = masked.  Built on ``device`` with
    one slice-assign per query frame, exactly like the reference (on a GPU that is x_len tiny kernel launches per block)."""
    ratio = c_len / x_len
    half = ratio / 2 + pad
    mask = torch.ones(batch, x_len, c_len, dtype=torch.bool, device=device)
    for i in range(x_len):
        mid = (i + 0.5) * ratio
        lo = max(round(mid - half), 0)
        hi = min(round(mid + half), c_len)
        mask[:, i, lo:hi] = False
    return mask


def _attention(sd: SD, p: str, x: torch.Tensor, context: Optional[torch.Tensor], mask: Optional[torch.Tensor], heads: int) -> torch.Tensor:
    """``CrossAttention.forward`` (``attention.py:86-128``)."""
    ctx = x if context is None else context
    q = F.linear(x, sd[p + "to_q.weight"])
    k = F.linear(ctx, sd[p + "to_k.weight"])
    v = F.linear(ctx, sd[p + "to_v.weight"])
    B, N, C = q.shape
    d = C // heads

    def split(t: torch.Tensor) -> torch.Tensor:  # 'b n (h d) -> (b h) n d'
        return t.reshape(B, t.shape[1], heads, d).permute(0, 2, 1, 3).reshape(B * heads, t.shape[1], d)

    q, k, v = split(q), split(k), split(v)
    sim = torch.einsum("bid,bjd->bij", q, k) * (d**-0.5)
    if mask is not None:
        m = mask[:, None].expand(B, heads, *mask.shape[1:]).reshape(B * heads, *mask.shape[1:])
        sim = sim.masked_fill(m, -torch.finfo(sim.dtype).max)
    attn = sim.softmax(dim=-1)
    out = torch.einsum("bij,bjd->bid", attn, v)
    out = out.reshape(B, heads, N, d).permute(0, 2, 1, 3).reshape(B, N, C)
    return F.linear(out, sd[p + "to_out.0.weight"], sd[p + "to_out.0.bias"])


def _transformer(sd: SD, p: str, x: torch.Tensor, context: torch.Tensor, heads: int = 6, taps: Optional[dict] = None) -> torch.Tensor:
    """``SpatialTransformer.forward`` + ``BasicTransformerBlock._forward`` (``attention.py:223-234, 167-193``)."""
    x_in = x
    x = F.group_norm(x, 32, sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-6)
    x = x.transpose(1, 2)  # b c t -> b t c
    tb = p + "transformer_blocks.0."
    c = x.shape[-1]
    x = _attention(sd, tb + "attn1.", F.layer_norm(x, (c,), sd[tb + "norm1.weight"], sd[tb + "norm1.bias"], 1e-5), None, None, heads) + x
    if taps is not None:
        taps[p + "after_attn1"] = x
    mask = alignment_mask(x.shape[0], x.shape[1], context.shape[1], device=x.device)
    x = _attention(sd, tb + "attn2.", F.layer_norm(x, (c,), sd[tb + "norm2.weight"], sd[tb + "norm2.bias"], 1e-5), context, mask, heads) + x
    if taps is not None:
        taps[p + "after_attn2"] = x
    h = F.layer_norm(x, (c,), sd[tb + "norm3.weight"], sd[tb + "norm3.bias"], 1e-5)
    h = F.linear(h, sd[tb + "ff.net.0.proj.weight"], sd[tb + "ff.net.0.proj.bias"])
    val, gate = h.chunk(2, dim=-1)
    h = val * F.gelu(gate)
    x = F.linear(h, sd[tb + "ff.net.2.weight"], sd[tb + "ff.net.2.bias"]) + x
    if taps is not None:
        taps[p + "after_ff"] = x
    x = x.transpose(1, 2)
    x = F.conv1d(x, sd[p + "proj_out.weight"], sd[p + "proj_out.bias"])
    return x + x_in


def denoiser_forward(sd: SD, sample: torch.Tensor, timesteps: torch.Tensor, context: torch.Tensor, prefix: str = "denoiser.model.", taps: Optional[dict] = None) -> torch.Tensor:
    """``UNet1DConditionModel.forward`` (``unet_1d_condition.py:51-77``) -> ``UNetModel.forward``
    (``openaimodel.py:677-709``).  ``sample`` (B,T,C_in), ``timesteps`` (B,), ``context`` (B,T_ctx,D)
    -> (B,T,C_in).  ``taps`` (optional dict) receives named intermediate activations in (B,C,T) /
    (B,T,C) layout for per-block parity tests."""
    g = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
    dt = sample.dtype
    x = sample.transpose(1, 2)
    emb = timestep_embedding(timesteps, 192, dt).to(dt)
    emb = F.linear(emb, g["time_embed.0.weight"], g["time_embed.0.bias"])
    emb = F.linear(F.silu(emb), g["time_embed.2.weight"], g["time_embed.2.bias"])
    if taps is not None:
        taps["emb"] = emb

    def tap(name: str, v: torch.Tensor) -> torch.Tensor:
        if taps is not None:
            taps[name] = v
        return v

    h0 = tap("input_blocks.0", F.conv1d(x, g["input_blocks.0.0.weight"], g["input_blocks.0.0.bias"], padding=1))
    h = tap("input_blocks.1.0", _resblock(g, "input_blocks.1.0.", h0, emb))
    h1 = tap("input_blocks.1.1", _transformer(g, "input_blocks.1.1.", h, context, taps=taps))
    h = tap("middle_block.0", _resblock(g, "middle_block.0.", h1, emb))
    h = tap("middle_block.1", _transformer(g, "middle_block.1.", h, context, taps=taps))
    h = tap("middle_block.2", _resblock(g, "middle_block.2.", h, emb))
    h = tap("output_blocks.0.0", _resblock(g, "output_blocks.0.0.", torch.cat([h, h1], dim=1), emb))
    h = tap("output_blocks.0.1", _transformer(g, "output_blocks.0.1.", h, context, taps=taps))
    h = tap("output_blocks.1.0", _resblock(g, "output_blocks.1.0.", torch.cat([h, h0], dim=1), emb))
    h = tap("output_blocks.1.1", _transformer(g, "output_blocks.1.1.", h, context, taps=taps))
    h = F.silu(_group_norm(h, g["out.0.weight"], g["out.0.bias"], 1e-5))
    out = F.conv1d(h, g["out.2.weight"], g["out.2.bias"], padding=1)
    return out.transpose(1, 2)


# --------------------------------------------------------------------------------------------------
# Audio encoder (said/model/wav2vec2.py over transformers' Wav2Vec2Model, base configuration)
# --------------------------------------------------------------------------------------------------

CONV_STRIDES = (5, 2, 2, 2, 2, 2, 2)


def folded_pos_conv_weight(sd: SD, p: str) -> torch.Tensor:
    """Weight-norm fold ``w = g * v / ||v||`` with the norm over dims (0, 1) per kernel tap
    (``nn.utils.weight_norm(conv, name="weight", dim=2)``, transformers ``Wav2Vec2PositionalConvEmbedding``)."""
    gk, vk = p + "weight_g", p + "weight_v"
    if gk not in sd:
        gk, vk = p + "parametrizations.weight.original0", p + "parametrizations.weight.original1"
    g, v = sd[gk], sd[vk]
    return g * v / v.norm(p=2, dim=(0, 1), keepdim=True)


def wav2vec2_forward(sd: SD, input_values: torch.Tensor, num_frames: Optional[int], prefix: str = "audio_encoder.", heads: Optional[int] = None, taps: Optional[dict] = None, stable_layer_norm: bool = False) -> torch.Tensor:
    """``ModifiedWav2Vec2Model.forward(...).last_hidden_state`` (``said/model/wav2vec2.py:14-82``) in eval
    mode without attention mask: conv feature encoder -> linear interpolation to ``num_frames``
    (``align_corners=True``) -> LN + projection -> grouped positional conv -> post-LN layers (wav2vec2-base family) or,
    with ``stable_layer_norm`` (wav2vec2-large family: ``do_stable_layer_norm=True``, TF modeling_wav2vec2.py:612-655,
    730-803), pre-LN layers and one LayerNorm after the last layer.  A feature extractor whose state dict holds a
    LayerNorm per conv layer is the ``feat_extract_norm="layer"`` variant (conv bias + LayerNorm over channels + GELU,
    TF :275-299).  ``heads`` defaults to hidden / 64."""
    g = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
    x = input_values[:, None]
    n_conv = sum(1 for k in g if k.startswith("feature_extractor.conv_layers.") and k.endswith("conv.weight"))
    for i in range(n_conv):
        w = g[f"feature_extractor.conv_layers.{i}.conv.weight"]
        layer_mode = "feature_extractor.conv_layers.1.layer_norm.weight" in g
        x = F.conv1d(x, w, g.get(f"feature_extractor.conv_layers.{i}.conv.bias"), stride=CONV_STRIDES[i])
        if layer_mode:
            x = F.layer_norm(x.transpose(-2, -1), (w.shape[0],), g[f"feature_extractor.conv_layers.{i}.layer_norm.weight"],
                             g[f"feature_extractor.conv_layers.{i}.layer_norm.bias"], 1e-5).transpose(-2, -1)
        elif i == 0:
            x = F.group_norm(x, w.shape[0], g["feature_extractor.conv_layers.0.layer_norm.weight"], g["feature_extractor.conv_layers.0.layer_norm.bias"], 1e-5)
        x = F.gelu(x)
        if taps is not None:
            taps[f"conv{i}"] = x
    if num_frames is not None:
        x = F.interpolate(x, size=num_frames, align_corners=True, mode="linear")
    x = x.transpose(1, 2)
    c = x.shape[-1]
    x = F.layer_norm(x, (c,), g["feature_projection.layer_norm.weight"], g["feature_projection.layer_norm.bias"], 1e-5)
    x = F.linear(x, g["feature_projection.projection.weight"], g["feature_projection.projection.bias"])
    if taps is not None:
        taps["projected"] = x
    hdim = x.shape[-1]
    wpos = folded_pos_conv_weight(g, "encoder.pos_conv_embed.conv.")
    k = wpos.shape[-1]
    pos = F.conv1d(x.transpose(1, 2), wpos, g["encoder.pos_conv_embed.conv.bias"], padding=k // 2, groups=hdim // wpos.shape[1])
    if k % 2 == 0:
        pos = pos[:, :, :-1]
    x = x + F.gelu(pos).transpose(1, 2)
    if not stable_layer_norm:
        x = F.layer_norm(x, (hdim,), g["encoder.layer_norm.weight"], g["encoder.layer_norm.bias"], 1e-5)
    if taps is not None:
        taps["encoder_in"] = x
    n_layers = sum(1 for kk in g if kk.endswith("final_layer_norm.weight"))
    if heads is None:
        heads = hdim // 64
    d = hdim // heads
    B, T, _ = x.shape
    sh = lambda t: t.reshape(B, T, heads, d).transpose(1, 2)  # noqa: E731

    def attention(p: str, h: torch.Tensor) -> torch.Tensor:
        q = F.linear(h, g[p + "attention.q_proj.weight"], g[p + "attention.q_proj.bias"]) * (d**-0.5)
        kk = F.linear(h, g[p + "attention.k_proj.weight"], g[p + "attention.k_proj.bias"])
        v = F.linear(h, g[p + "attention.v_proj.weight"], g[p + "attention.v_proj.bias"])
        att = torch.softmax(sh(q) @ sh(kk).transpose(-1, -2), dim=-1) @ sh(v)
        att = att.transpose(1, 2).reshape(B, T, hdim)
        return F.linear(att, g[p + "attention.out_proj.weight"], g[p + "attention.out_proj.bias"])

    if stable_layer_norm:
        for l in range(n_layers):
            p = f"encoder.layers.{l}."
            x = x + attention(p, F.layer_norm(x, (hdim,), g[p + "layer_norm.weight"], g[p + "layer_norm.bias"], 1e-5))
            h = F.layer_norm(x, (hdim,), g[p + "final_layer_norm.weight"], g[p + "final_layer_norm.bias"], 1e-5)
            ff = F.gelu(F.linear(h, g[p + "feed_forward.intermediate_dense.weight"], g[p + "feed_forward.intermediate_dense.bias"]))
            x = x + F.linear(ff, g[p + "feed_forward.output_dense.weight"], g[p + "feed_forward.output_dense.bias"])
            if taps is not None:
                taps[f"layer{l}"] = x
        return F.layer_norm(x, (hdim,), g["encoder.layer_norm.weight"], g["encoder.layer_norm.bias"], 1e-5)
    for l in range(n_layers):
        p = f"encoder.layers.{l}."
        att = attention(p, x)
        x = F.layer_norm(x + att, (hdim,), g[p + "layer_norm.weight"], g[p + "layer_norm.bias"], 1e-5)
        ff = F.gelu(F.linear(x, g[p + "feed_forward.intermediate_dense.weight"], g[p + "feed_forward.intermediate_dense.bias"]))
        ff = F.linear(ff, g[p + "feed_forward.output_dense.weight"], g[p + "feed_forward.output_dense.bias"])
        x = F.layer_norm(x + ff, (hdim,), g[p + "final_layer_norm.weight"], g[p + "final_layer_norm.bias"], 1e-5)
        if taps is not None:
            taps[f"layer{l}"] = x
    return x


def audio_embedding(sd: SD, waveform: torch.Tensor, num_frames: int) -> torch.Tensor:
    """``SAID.get_audio_embedding`` (``diffusion.py:209-230``)."""
    feats = wav2vec2_forward(sd, waveform, num_frames)
    if "audio_proj_layer.weight" in sd:
        feats = F.linear(feats, sd["audio_proj_layer.weight"], sd["audio_proj_layer.bias"])
    return feats


# --------------------------------------------------------------------------------------------------
# The pipeline (said/model/diffusion.py:308-472)
# --------------------------------------------------------------------------------------------------


def process_audio(waveform: np.ndarray) -> torch.Tensor:
    """``SAID.process_audio`` (``diffusion.py:188-207``) for one utterance: zero-mean unit-variance."""
    x = np.asarray(waveform, dtype=np.float32)
    return torch.from_numpy(((x - x.mean()) / np.sqrt(x.var() + 1e-7)).astype(np.float32))[None]


def inference(
    sd: SD,
    waveform_processed: torch.Tensor,
    init_samples: Optional[torch.Tensor] = None,
    mask: Optional[torch.Tensor] = None,
    num_inference_steps: int = 100,
    strength: float = 1.0,
    guidance_scale: float = 2.5,
    guidance_rescale: float = 0.0,
    eta: float = 0.0,
    fps: int = 60,
    save_intermediate: bool = False,
    prediction_type: str = "epsilon",
    latent_scale: float = 1.0,
    sampling_rate: int = 16000,
    noise: Optional[torch.Tensor] = None,
    eta_noise: Optional[torch.Tensor] = None,
    audio_emb: Optional[torch.Tensor] = None,
    teacher_latents: Optional[List[torch.Tensor]] = None,
    return_preclamp: bool = False,
    step_limit: Optional[int] = None,
    scheduler: str = "ddim",
) -> Tuple[torch.Tensor, List[torch.Tensor]]:
    """``SAID.inference`` restated.  ``noise`` replaces the single ``torch.randn`` draw of
    ``diffusion.py:364`` (generation) or ``:270`` (editing) so that the CUDA path can be fed the very same
    tensor; ``eta_noise`` (n_steps,B,T,C) replaces the per-step draws for ``eta > 0``.  ``teacher_latents``
    (list, one per step) overrides the latents fed to each step (teacher forcing).  ``step_limit`` stops
    after that many loop iterations (bounded CPU-baseline sample).  Returns ``(result, intermediates)``."""
    dt = waveform_processed.dtype
    B, T_a = waveform_processed.shape
    C = sd["denoiser.model.out.2.bias"].shape[0]
    do_cfg = guidance_scale > 1.0
    T = int(T_a / sampling_rate * fps)
    ac = ddim_alphas_cumprod(1000)
    timesteps = ddim_timesteps(num_inference_steps)

    if init_samples is None:
        latents = noise.clone().to(dt) if noise is not None else torch.randn(B, T, C, dtype=torch.float32).to(dt)
    else:
        latents = init_samples.clone().to(dt)
    latents = latents * (latent_scale * 1.0)
    init_latents = latents.clone()
    init_timestep = min(int(num_inference_steps * strength), num_inference_steps)
    edit_noise = None
    if init_samples is not None:
        t_star = int(timesteps[-init_timestep])
        edit_noise = noise.clone().to(dt) if noise is not None else torch.randn(latents.shape, dtype=torch.float32).to(dt)
        latents = ddim_add_noise(latents, edit_noise, [t_star] * B, ac)

    emb = audio_embedding(sd, waveform_processed, T) if audio_emb is None else audio_emb
    if do_cfg:
        uncond = sd["null_cond_emb"].to(dt).repeat(B, emb.shape[1], 1)
        emb = torch.cat([uncond, emb])

    intermediates: List[torch.Tensor] = []
    t_start = num_inference_steps - init_timestep
    for idx, t in enumerate(timesteps[t_start:]):
        if step_limit is not None and idx >= step_limit:
            break
        if teacher_latents is not None:
            latents = teacher_latents[idx].to(dt)
        if save_intermediate:
            intermediates.append((latents / latent_scale).clone())
        x_in = torch.cat([latents] * 2) if do_cfg else latents
        ts = torch.full((x_in.shape[0],), int(t), dtype=torch.long, device=x_in.device)
        pred = denoiser_forward(sd, x_in, ts, emb)
        if do_cfg:
            u, c = pred.chunk(2)
            pred = c + guidance_scale * (c - u)
            if guidance_rescale > 0.0:
                pred = rescale_noise_cfg(pred, c, guidance_rescale)
        if scheduler == "ddpm":   # eta_noise then holds DDPMScheduler's per-step variance noise (row unused at t = 0)
            latents = ddpm_step(pred, int(t), latents, ac, num_inference_steps, prediction_type,
                                variance_noise=None if eta_noise is None else eta_noise[idx].to(dt))
        else:
            latents = ddim_step(
                pred, int(t), latents, ac, num_inference_steps, prediction_type, eta,
                variance_noise=None if eta_noise is None else eta_noise[idx].to(dt),
            )
        if init_samples is not None and mask is not None:
            noisy = init_latents
            nxt = t_start + idx + 1
            if nxt < num_inference_steps:
                noisy = ddim_add_noise(init_latents, edit_noise, [int(timesteps[nxt])], ac)
            latents = noisy * mask + latents * (1 - mask)
    pre = latents / latent_scale
    if return_preclamp:
        intermediates.append(pre.clone())
    return pre.clone().clamp(0, 1), intermediates
