"""ctypes binding of ``libsaid_sm100.so`` (C ABI declared in ``include/said_b200.h``).

This is the only place Python touches the native library.  There is no fallback: if the shared
library is missing or the device is not sm_100, the product path raises.
"""
from __future__ import annotations

import ctypes
import os
from typing import Dict, Optional, Sequence

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsaid_sm100.so")

# every symbol include/said_b200.h declares (tests check the .so exports exactly these)
EXPORTED_SYMBOLS = (
    "said_last_error",
    "said_version",
    "said_create",
    "said_destroy",
    "said_set_tensor",
    "said_commit_weights",
    "said_weights_ready",
    "said_get_config",
    "said_encode_audio",
    "said_normalize_audio",
    "said_resample_mono",
    "said_prepare_context",
    "said_denoise",
    "said_denoiser_forward",
    "said_check_status",
    "said_eval_commit_bcvae",
    "said_eval_bcvae_latents",
    "said_eval_frechet",
    "said_op_gemm_h",
    "said_op_ffn_h",
    "said_op_gemm_h_bench",
    "said_op_ddim_step",
    "said_op_self_attention",
    "said_op_self_attention_tc",
    "said_op_self_attention_h",
    "said_launch_count",
    "said_graph_captures",
    "said_op_gemm_tc_bench",
    "said_set_precision",
    "said_profile_begin",
    "said_profile_end",
)


class DenoiseArgs(ctypes.Structure):
    """``said_denoise_args`` (include/said_b200.h)."""

    _fields_ = [
        ("B", ctypes.c_int),
        ("T", ctypes.c_int),
        ("n_steps", ctypes.c_int),
        ("timesteps_host", ctypes.c_void_p),
        ("step_table_host", ctypes.c_void_p),
        ("prediction_type", ctypes.c_int),
        ("do_cfg", ctypes.c_int),
        ("guidance_scale", ctypes.c_float),
        ("guidance_rescale", ctypes.c_float),
        ("latent_scale", ctypes.c_float),
        ("init_src_dev", ctypes.c_void_p),
        ("init_scale", ctypes.c_float),
        ("edit_noise_dev", ctypes.c_void_p),
        ("edit_sqrt_a", ctypes.c_float),
        ("edit_sqrt_b", ctypes.c_float),
        ("mask_dev", ctypes.c_void_p),
        ("eta_noise_dev", ctypes.c_void_p),
        ("intermediates_dev", ctypes.c_void_p),
        ("result_dev", ctypes.c_void_p),
        ("latents_out_dev", ctypes.c_void_p),
        ("use_graph", ctypes.c_int),
        ("scheduler", ctypes.c_int),
        ("resume", ctypes.c_int),
        ("more", ctypes.c_int),
    ]


class SaidLibraryError(RuntimeError):
    pass


_lib: Optional[ctypes.CDLL] = None


def load_library() -> ctypes.CDLL:
    """Load the native library or raise -- never falls back to anything else."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SaidLibraryError(
            f"{LIB_PATH} is not built. Build it with `make -C said_b200/csrc` "
            "(or `python -c 'import __graft_entry__ as g; g.build()'`); said_b200 has no non-CUDA path."
        )
    lib = ctypes.CDLL(LIB_PATH)
    vp, ci, cf = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
    lib.said_last_error.restype = ctypes.c_char_p
    lib.said_last_error.argtypes = []
    lib.said_version.restype = ci
    lib.said_create.argtypes = [ci, ctypes.POINTER(vp)]
    lib.said_destroy.argtypes = [vp]
    lib.said_destroy.restype = None
    lib.said_set_tensor.argtypes = [vp, ctypes.c_char_p, vp, ctypes.POINTER(ctypes.c_int64), ci]
    lib.said_commit_weights.argtypes = [vp]
    lib.said_weights_ready.argtypes = [vp]
    lib.said_get_config.argtypes = [vp, ctypes.POINTER(ci), ctypes.POINTER(ci), ctypes.POINTER(ci)]
    lib.said_encode_audio.argtypes = [vp, vp, ci, ci, ci, vp, vp]
    lib.said_normalize_audio.argtypes = [vp, vp, ci, ci, vp, vp]
    lib.said_resample_mono.argtypes = [vp, vp, ci, ci, ci, ci, ci, vp, vp, ci, vp]
    lib.said_prepare_context.argtypes = [vp, vp, ci, ci, ci, vp]
    lib.said_denoise.argtypes = [vp, ctypes.POINTER(DenoiseArgs), vp]
    lib.said_denoiser_forward.argtypes = [vp, vp, vp, vp, ci, ci, ci, vp, vp, vp]
    lib.said_check_status.argtypes = [vp, vp, ctypes.POINTER(ci)]
    lib.said_eval_commit_bcvae.argtypes = [vp]
    lib.said_eval_bcvae_latents.argtypes = [vp, vp, ci, ci, ci, vp, vp]
    lib.said_eval_frechet.argtypes = [vp, vp, ci, vp, ci, ctypes.POINTER(ctypes.c_double), vp]
    lib.said_op_gemm_h.argtypes = [vp, vp, ci, ci, ci, vp, ci, vp, vp, vp]
    lib.said_op_ffn_h.argtypes = [vp, vp, vp, vp, ci, vp, vp, vp, vp, vp, vp]
    lib.said_op_gemm_h_bench.argtypes = [vp, ci, ci, ci, ci, ci, ci, ci, ctypes.POINTER(cf)]
    lib.said_op_ddim_step.argtypes = [vp, vp, vp, ci, ci, ci, cf, cf, ci, vp, vp, ci, vp]
    lib.said_op_self_attention.argtypes = [vp, vp, ci, ci, ci, ci, vp, vp]
    lib.said_op_self_attention_tc.argtypes = [vp, vp, ci, ci, ci, vp, vp]
    lib.said_op_self_attention_h.argtypes = [vp, vp, ci, ci, ci, vp, vp]
    lib.said_launch_count.argtypes = [vp]
    lib.said_launch_count.restype = ctypes.c_longlong
    lib.said_graph_captures.argtypes = [vp]
    lib.said_graph_captures.restype = ctypes.c_longlong
    lib.said_set_precision.argtypes = [vp, ci, ci, ci]
    lib.said_op_gemm_tc_bench.argtypes = [vp, ci, ci, ci, ci, ci, ci, ctypes.POINTER(cf)]
    lib.said_profile_begin.argtypes = [vp]
    lib.said_profile_end.argtypes = [vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_longlong), ci]
    _lib = lib
    return lib


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _check_dev(t: torch.Tensor, device: torch.device, name: str) -> torch.Tensor:
    if t.device != device:
        raise ValueError(f"{name} is on {t.device}, the engine runs on {device}")
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {t.dtype}")
    return t.contiguous()


class Engine:
    """One native engine on one CUDA device."""

    def __init__(self, device: torch.device):
        device = torch.device(device)
        if device.type != "cuda":
            raise SaidLibraryError(f"said_b200 runs on CUDA sm_100a devices only (got device '{device}')")
        self.lib = load_library()
        self.device = torch.device("cuda", device.index if device.index is not None else torch.cuda.current_device())
        h = ctypes.c_void_p()
        self._h = None
        self._call(self.lib.said_create(self.device.index, ctypes.byref(h)))
        self._h = h
        self._keep: list = []

    def _call(self, rc: int) -> None:
        if rc != 0:
            raise SaidLibraryError(self.lib.said_last_error().decode("utf-8", "replace"))

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self.lib.said_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    # ------------------------------------------------------------------ weights
    def set_tensor(self, name: str, value: torch.Tensor) -> None:
        a = np.ascontiguousarray(value.detach().to("cpu", torch.float32).numpy())
        shape = (ctypes.c_int64 * a.ndim)(*a.shape)
        self._call(self.lib.said_set_tensor(self._h, name.encode(), a.ctypes.data, shape, a.ndim))

    def load_weights(self, tensors: Dict[str, torch.Tensor]) -> None:
        for k, v in tensors.items():
            self.set_tensor(k, v)
        with torch.cuda.device(self.device):
            self._call(self.lib.said_commit_weights(self._h))

    def config(self):
        a, b, c = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        self._call(self.lib.said_get_config(self._h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)))
        return {"in_channels": a.value, "ctx_dim": b.value, "enc_hidden": c.value}

    @property
    def launches(self) -> int:
        return int(self.lib.said_launch_count(self._h))

    @property
    def graph_captures(self) -> int:
        return int(self.lib.said_graph_captures(self._h))

    PROFILE_FAMILIES = ("gemm_conv3", "gemm_layernorm", "gemm_plain", "self_attention", "cross_attention3",
                        "gn_stats", "cfg_ddim_step", "other")

    PRECISIONS = {"fp32": 0, "tf32x3": 1, "tf32": 2, "fp16x3": 3}

    def set_precision(self, mode: str, tc_min_rows: int = 0, encoder_mode: str = "fp32") -> None:
        if mode not in self.PRECISIONS or encoder_mode not in self.PRECISIONS:
            raise ValueError(f"precision must be one of {list(self.PRECISIONS)}")
        self._call(self.lib.said_set_precision(self._h, self.PRECISIONS[mode], int(tc_min_rows), self.PRECISIONS[encoder_mode]))

    def profile_begin(self) -> None:
        self._call(self.lib.said_profile_begin(self._h))

    def profile_end(self) -> Dict[str, Dict[str, float]]:
        n = len(self.PROFILE_FAMILIES)
        ms = (ctypes.c_double * n)()
        cnt = (ctypes.c_longlong * n)()
        self._call(self.lib.said_profile_end(self._h, ms, cnt, n))
        return {k: {"ms": float(ms[i]), "launches": int(cnt[i])} for i, k in enumerate(self.PROFILE_FAMILIES)}

    # ------------------------------------------------------------------ phases
    def encode_audio(self, wave: torch.Tensor, num_frames: int) -> torch.Tensor:
        wave = _check_dev(wave, self.device, "waveform")
        B, T_a = wave.shape
        out = torch.empty((B, num_frames, self.config()["ctx_dim"]), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            self._call(self.lib.said_encode_audio(self._h, wave.data_ptr(), B, T_a, num_frames, out.data_ptr(), self._stream()))
        return out

    def normalize_audio(self, wave: torch.Tensor) -> torch.Tensor:
        wave = _check_dev(wave, self.device, "waveform")
        B, T_a = wave.shape
        out = torch.empty_like(wave)
        with torch.cuda.device(self.device):
            self._call(self.lib.said_normalize_audio(self._h, wave.data_ptr(), B, T_a, out.data_ptr(), self._stream()))
        return out

    def resample_mono(self, wave: torch.Tensor, orig_freq: int, new_freq: int) -> torch.Tensor:
        """(channels, n) waveform on the device at ``orig_freq`` -> mono (n_out,) at ``new_freq``: torchaudio.functional.resample
        (sinc_interp_hann, lowpass_filter_width 6, rolloff 0.99 -- its defaults, which the reference uses) + mean over channels."""
        import math

        wave = _check_dev(wave, self.device, "waveform")
        channels, n_in = wave.shape
        g = math.gcd(int(orig_freq), int(new_freq))
        orig, nw = int(orig_freq) // g, int(new_freq) // g
        # filter bank: torchaudio/functional/functional.py::_get_sinc_resample_kernel, evaluated in float32 as resample() does
        lpw, rolloff = 6, 0.99
        base = min(orig, nw) * rolloff
        width = math.ceil(lpw * orig / base)
        idx = torch.arange(-width, width + orig, dtype=torch.float32)[None] / orig
        t = torch.arange(0, -nw, -1, dtype=torch.float32)[:, None] / nw + idx
        t = (t * base).clamp_(-lpw, lpw)
        window = torch.cos(t * math.pi / lpw / 2) ** 2
        t = t * math.pi
        bank = torch.where(t == 0, torch.tensor(1.0), t.sin() / t) * window * (base / orig)
        bank = bank.to(torch.float32).contiguous().to(self.device)
        n_out = int(math.ceil(nw * n_in / orig))
        out = torch.empty((n_out,), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            self._call(self.lib.said_resample_mono(self._h, wave.data_ptr(), channels, n_in, orig, nw, width, bank.data_ptr(),
                                                   out.data_ptr(), n_out, self._stream()))
        self._keep_aux = [wave, bank]
        return out

    def prepare_context(self, emb: torch.Tensor, with_uncond: bool) -> None:
        emb = _check_dev(emb, self.device, "audio embedding")
        B, T, _ = emb.shape
        self._keep = [emb]
        with torch.cuda.device(self.device):
            self._call(self.lib.said_prepare_context(self._h, emb.data_ptr(), B, T, int(with_uncond), self._stream()))

    def denoise(
        self,
        init_src: torch.Tensor,
        timesteps: Sequence[int],
        step_table: np.ndarray,
        prediction_type: int,
        do_cfg: bool,
        guidance_scale: float,
        guidance_rescale: float,
        latent_scale: float,
        init_scale: float,
        edit_noise: Optional[torch.Tensor] = None,
        edit_coefs=(1.0, 0.0),
        mask: Optional[torch.Tensor] = None,
        eta_noise: Optional[torch.Tensor] = None,
        intermediates: Optional[torch.Tensor] = None,
        latents_out: Optional[torch.Tensor] = None,
        use_graph: bool = True,
        scheduler: int = 0,
        resume: bool = False,
        more: bool = False,
    ) -> torch.Tensor:
        init_src = _check_dev(init_src, self.device, "initial latents")
        B, T, C = init_src.shape
        n_steps = len(timesteps)
        ts = np.ascontiguousarray(np.asarray(timesteps, dtype=np.float32))
        tab = np.ascontiguousarray(step_table, dtype=np.float32).reshape(n_steps, 8)
        result = torch.empty((B, T, C), dtype=torch.float32, device=self.device)
        opt = {}
        for name, t in (("edit_noise", edit_noise), ("mask", mask), ("eta_noise", eta_noise)):
            opt[name] = None if t is None else _check_dev(t, self.device, name)
        for name, t in (("intermediates", intermediates), ("latents_out", latents_out)):
            if t is not None and (t.device != self.device or t.dtype != torch.float32 or not t.is_contiguous()):
                raise ValueError(f"{name} must be a contiguous float32 tensor on {self.device}")
        a = DenoiseArgs(
            B=B, T=T, n_steps=n_steps,
            timesteps_host=ts.ctypes.data if n_steps else None,
            step_table_host=tab.ctypes.data if n_steps else None,
            prediction_type=int(prediction_type), do_cfg=int(bool(do_cfg)),
            guidance_scale=float(guidance_scale), guidance_rescale=float(guidance_rescale),
            latent_scale=float(latent_scale),
            init_src_dev=init_src.data_ptr(), init_scale=float(init_scale),
            edit_noise_dev=_ptr(opt["edit_noise"]), edit_sqrt_a=float(edit_coefs[0]), edit_sqrt_b=float(edit_coefs[1]),
            mask_dev=_ptr(opt["mask"]), eta_noise_dev=_ptr(opt["eta_noise"]),
            intermediates_dev=_ptr(intermediates), result_dev=result.data_ptr(),
            latents_out_dev=_ptr(latents_out), use_graph=int(bool(use_graph)), scheduler=int(scheduler),
            resume=int(bool(resume)), more=int(bool(more)),
        )
        with torch.cuda.device(self.device):
            self._call(self.lib.said_denoise(self._h, ctypes.byref(a), self._stream()))
        # the kernels read these asynchronously; keep them alive until the caller synchronises / next call
        self._keep = [init_src, opt, intermediates, latents_out, result] + self._keep[:1]
        return result

    def denoiser_forward(self, x: torch.Tensor, timesteps: torch.Tensor, ctx: torch.Tensor, taps: bool = False):
        x = _check_dev(x, self.device, "noisy samples")
        ctx = _check_dev(ctx, self.device, "audio embedding")
        Bp, T, C = x.shape
        if ctx.shape[0] != Bp:
            raise ValueError(f"audio embedding must have batch {Bp}; got {tuple(ctx.shape)}")
        T_ctx = int(ctx.shape[1])
        ts = np.ascontiguousarray(timesteps.detach().to("cpu").reshape(-1).numpy().astype(np.float32))
        if ts.shape[0] == 1 and Bp > 1:
            ts = np.repeat(ts, Bp)
        if ts.shape[0] != Bp:
            raise ValueError(f"timesteps must have 1 or {Bp} entries, got {ts.shape[0]}")
        out = torch.empty_like(x)
        tap_buf = torch.empty((10, Bp, T, 192), dtype=torch.float32, device=self.device) if taps else None
        with torch.cuda.device(self.device):
            self._call(self.lib.said_denoiser_forward(self._h, x.data_ptr(), ts.ctypes.data, ctx.data_ptr(), Bp, T, T_ctx,
                                                      out.data_ptr(), _ptr(tap_buf), self._stream()))
        self.check_status()
        return (out, tap_buf) if taps else out

    # ------------------------------------------------------------------ evaluation
    def load_bcvae(self, state_dict: Dict[str, torch.Tensor]) -> None:
        """Encoder half of the reference's ``BCVAE`` state dict (``said/model/vae.py``; e.g. ``torch.load("model/vae.pth")``)."""
        for k, v in state_dict.items():
            if k.startswith("encoder.") and "num_batches_tracked" not in k and "fc_logvar" not in k:
                self.set_tensor("bcvae." + k, v)
        with torch.cuda.device(self.device):
            self._call(self.lib.said_eval_commit_bcvae(self._h))

    def bcvae_latents(self, coeffs: torch.Tensor, step: int) -> torch.Tensor:
        """(B, T, 32) coefficient sequences -> (B * nw, 64) BCVAE latent means over windows of 120 frames every `step` frames."""
        coeffs = _check_dev(coeffs, self.device, "coefficients")
        B, T, C = coeffs.shape
        if C != 32 or T < 120:
            raise ValueError("bcvae_latents: coefficients must be (B, T >= 120, 32)")
        nw = (T - 120) // int(step) + 1
        out = torch.empty((B * nw, 64), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            self._call(self.lib.said_eval_bcvae_latents(self._h, coeffs.data_ptr(), B, T, int(step), out.data_ptr(), self._stream()))
        return out

    def frechet(self, lat1: torch.Tensor, lat2: torch.Tensor) -> Dict[str, float]:
        lat1 = _check_dev(lat1, self.device, "latents 1")
        lat2 = _check_dev(lat2, self.device, "latents 2")
        out = (ctypes.c_double * 4)()
        with torch.cuda.device(self.device):
            self._call(self.lib.said_eval_frechet(self._h, lat1.data_ptr(), lat1.shape[0], lat2.data_ptr(), lat2.shape[0], out, self._stream()))
        return {"frechet_distance": out[0], "mean_term": out[1], "trace_term": out[2], "sqrt_term": out[3]}

    def check_status(self) -> None:
        """Synchronise and raise if the device-side status word is set (fp16x3 path: an activation left fp16's range)."""
        v = ctypes.c_int(0)
        with torch.cuda.device(self.device):
            self._call(self.lib.said_check_status(self._h, self._stream(), ctypes.byref(v)))
        if v.value & 1:
            raise SaidLibraryError(
                "an activation reached fp16's range limit (|x| >= 65000) on the fp16x3 tensor-core path; the result of this call is "
                "invalid -- rerun with model.precision = 'tf32x3' (3xTF32 operands have fp32's range)"
            )

    # ------------------------------------------------------------------ unit ops (tests)
    def op_ddim_step(self, pred, latents, do_cfg, guidance_scale, guidance_rescale, prediction_type, row8, eta_noise=None, scheduler=0):
        pred = _check_dev(pred, self.device, "pred")
        latents = _check_dev(latents, self.device, "latents").clone()
        B = latents.shape[0]
        n = latents[0].numel()
        row = np.ascontiguousarray(row8, dtype=np.float32)
        en = None if eta_noise is None else _check_dev(eta_noise, self.device, "eta_noise")
        with torch.cuda.device(self.device):
            self._call(self.lib.said_op_ddim_step(self._h, pred.data_ptr(), latents.data_ptr(), B, n, int(do_cfg),
                                                  float(guidance_scale), float(guidance_rescale), int(prediction_type),
                                                  row.ctypes.data, _ptr(en), int(scheduler), self._stream()))
            torch.cuda.synchronize(self.device)
        return latents

    def op_self_attention_tc(self, qkv: torch.Tensor, heads: int) -> torch.Tensor:
        qkv = _check_dev(qkv, self.device, "qkv")
        B, T, W = qkv.shape
        assert W == 3 * heads * 32
        out = torch.empty((B, T, heads * 32), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            self._call(self.lib.said_op_self_attention_tc(self._h, qkv.data_ptr(), B, T, heads, out.data_ptr(), self._stream()))
        return out

    def op_self_attention_h(self, qkv: torch.Tensor, heads: int) -> torch.Tensor:
        qkv = _check_dev(qkv, self.device, "qkv")
        B, T, W = qkv.shape
        assert W == 3 * heads * 32
        out = torch.empty((B, T, heads * 32), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            self._call(self.lib.said_op_self_attention_h(self._h, qkv.data_ptr(), B, T, heads, out.data_ptr(), self._stream()))
        return out

    def op_gemm_h(self, a: torch.Tensor, wt: torch.Tensor, taps: int = 1, bias: Optional[torch.Tensor] = None) -> torch.Tensor:
        """fp16x3 GEMM unit op: a (M, Cin) on the device, wt (taps * Cin, N) K-major on the host -> (M, N)."""
        a = _check_dev(a, self.device, "a")
        M, Cin = a.shape
        w = np.ascontiguousarray(wt.detach().to("cpu", torch.float32).numpy())
        N = w.shape[1]
        assert w.shape[0] == taps * Cin
        b = None if bias is None else _check_dev(bias, self.device, "bias")
        out = torch.empty((M, N), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            self._call(self.lib.said_op_gemm_h(self._h, a.data_ptr(), M, Cin, taps, w.ctypes.data, N, _ptr(b), out.data_ptr(), self._stream()))
        return out

    def op_ffn_h(self, ln: torch.Tensor, x2: torch.Tensor, res: Optional[torch.Tensor], w1: torch.Tensor, b1: torch.Tensor,
                 w2: torch.Tensor, b2: Optional[torch.Tensor]) -> torch.Tensor:
        """fused feed-forward unit op (see said_op_ffn_h): ln, x2, res (M, 192) on the device; w1 (192, 1536) / w2 (960, 192) on the host."""
        ln = _check_dev(ln, self.device, "ln")
        x2 = _check_dev(x2, self.device, "x2")
        r = None if res is None else _check_dev(res, self.device, "res")
        M = ln.shape[0]
        w1h = np.ascontiguousarray(w1.detach().to("cpu", torch.float32).numpy())
        w2h = np.ascontiguousarray(w2.detach().to("cpu", torch.float32).numpy())
        assert w1h.shape == (192, 1536) and w2h.shape == (960, 192) and tuple(ln.shape) == (M, 192) == tuple(x2.shape)
        b1d = _check_dev(b1, self.device, "b1")
        b2d = None if b2 is None else _check_dev(b2, self.device, "b2")
        out = torch.empty((M, 192), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            self._call(self.lib.said_op_ffn_h(self._h, ln.data_ptr(), x2.data_ptr(), _ptr(r), M, w1h.ctypes.data, b1d.data_ptr(),
                                              w2h.ctypes.data, _ptr(b2d), out.data_ptr(), self._stream()))
        return out

    def op_gemm_h_bench(self, M: int, Cin: int, taps: int = 1, N: int = 192, with_residual: int = 1, dbg: int = 0, iters: int = 10) -> float:
        """with_residual: 0 plain store, 1 residual add, 2 the GEGLU epilogue with pair output"""
        ms = ctypes.c_float()
        with torch.cuda.device(self.device):
            self._call(self.lib.said_op_gemm_h_bench(self._h, M, Cin, taps, N, int(with_residual), dbg, iters, ctypes.byref(ms)))
        return float(ms.value)

    def op_gemm_tc_bench(self, M: int, K: int, nsplit: int = 3, with_residual: bool = True, dbg: int = 0, iters: int = 10) -> float:
        ms = ctypes.c_float()
        with torch.cuda.device(self.device):
            self._call(self.lib.said_op_gemm_tc_bench(self._h, M, K, nsplit, int(with_residual), dbg, iters, ctypes.byref(ms)))
        return float(ms.value)

    def op_self_attention(self, qkv: torch.Tensor, heads: int, head_dim: int) -> torch.Tensor:
        qkv = _check_dev(qkv, self.device, "qkv")
        B, T, W = qkv.shape
        assert W == 3 * heads * head_dim
        out = torch.empty((B, T, heads * head_dim), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            self._call(self.lib.said_op_self_attention(self._h, qkv.data_ptr(), B, T, heads, head_dim, out.data_ptr(), self._stream()))
        return out
