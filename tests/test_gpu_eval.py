"""GPU: on-device BCVAE window embedder and Frechet distance against the reference's own BCVAE class (golden) and the CPU oracle."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_bcvae_latents_golden(golden_dir):
    """Latent means of sliding windows == the reference's ``BCVAE.encode(window).mean`` (tests/golden/make_golden_bcvae.py)."""
    from oracle import eval_oracle as E
    from said_b200.evaluation import DeviceEvaluator

    gd = np.load(os.path.join(golden_dir, "bcvae_windows.npz"))
    ev = DeviceEvaluator(DEV, E.synthetic_bcvae_state_dict(0))
    lat = ev.latents(torch.from_numpy(gd["coeffs"]), int(gd["step"])).cpu()
    err = float((lat - torch.from_numpy(gd["latents"])).abs().max())
    print("bcvae latents vs reference", err, tuple(lat.shape))
    assert lat.shape == gd["latents"].shape
    assert err < 2e-5


def test_frechet_distance_vs_oracle():
    """Device Frechet distance (fp64 Jacobi) against the restated pytorch_fid formula (scipy sqrtm) on two latent clouds."""
    from oracle import eval_oracle as E
    from said_b200.evaluation import DeviceEvaluator

    ev = DeviceEvaluator(DEV, E.synthetic_bcvae_state_dict(0))
    g = torch.Generator().manual_seed(2)
    a = torch.randn(700, 64, generator=g) @ (0.3 * torch.randn(64, 64, generator=g)) + 0.2
    b = torch.randn(500, 64, generator=g) @ (0.25 * torch.randn(64, 64, generator=g)) - 0.1
    got = ev.engine.frechet(a.to(DEV), b.to(DEV))
    want = E.frechet_distance(a.double().numpy(), b.double().numpy())
    same = ev.engine.frechet(a.to(DEV), a.to(DEV))["frechet_distance"]
    print("frechet device", got["frechet_distance"], "oracle", want, "self-distance", same)
    assert abs(got["frechet_distance"] - want) < 1e-6 * abs(want)
    assert abs(same) < 1e-8 * got["trace_term"]


def test_frechet_between_precision_modes(gpu_model, state_dict):
    """The use the evaluator is built for: the OUTPUT DISTRIBUTION of the fp16x3 tensor-core path against the fp32 kernels on an
    epsilon-prediction chain (per-sample parity is meaningless there after many steps: the chain is chaotic).  The distance between
    the two precision modes must be a small fraction of the distance between two different noise seeds."""
    from oracle import eval_oracle as E
    from said_b200.evaluation import DeviceEvaluator
    from said_b200.synth import synthetic_batch

    m = gpu_model("epsilon")
    ev = DeviceEvaluator(DEV, E.synthetic_bcvae_state_dict(0))
    B, T = 24, 300
    wave = synthetic_batch(B, 5.0).to(DEV)
    g = torch.Generator().manual_seed(41)
    noise = torch.randn(2, B, T, 32, generator=g).to(DEV)
    outs = {}
    saved = m.precision
    try:
        for mode, k in (("fp16x3", 0), ("fp32", 0), ("fp16x3", 1)):
            m.precision = mode
            with torch.no_grad():
                outs[(mode, k)] = m._run(wave, noise[k], None, None, 100, 1.0, 2.0, 0.0, 0.0, T, False, False, None).result
    finally:
        m.precision = saved
    d_modes = ev.frechet_distance(outs[("fp16x3", 0)], outs[("fp32", 0)], 20)
    d_seeds = ev.frechet_distance(outs[("fp16x3", 0)], outs[("fp16x3", 1)], 20)
    print("frechet: fp16x3 vs fp32 (same noise)", d_modes, " different noise", d_seeds)
    assert d_modes < 0.05 * d_seeds


@pytest.mark.parametrize("sr,channels", [(44100, 2), (48000, 1), (8000, 2), (22050, 1)])
def test_resample_mono_vs_torchaudio(tmp_path, sr, channels):
    """Device-side load_audio tail (resample to 16 kHz + mono mix, 8(f) rank 1) against torchaudio.functional.resample +
    mean on the CPU -- what the reference's load_audio runs (said/util/audio.py:35-38)."""
    import torchaudio
    from scipy.io import wavfile

    from said_b200.util.audio import fit_audio_unet, load_audio, load_audio_device

    g = torch.Generator().manual_seed(sr + channels)
    n = int(sr * 1.3) + 17
    t = torch.arange(n, dtype=torch.float32) / sr
    x = torch.stack([0.4 * torch.sin(6.2832 * (200 + 50 * c) * t) + 0.05 * torch.randn(n, generator=g) for c in range(channels)])
    pcm = (x.clamp(-1, 1) * 32767).round().to(torch.int16)
    path = os.path.join(tmp_path, "a.wav")
    wavfile.write(path, sr, pcm.T.contiguous().numpy() if channels > 1 else pcm[0].numpy())
    want = load_audio(path, 16000)                      # host: torchaudio resample + mean
    got = load_audio_device(path, 16000, DEV)
    ref = torchaudio.functional.resample(pcm.float() / 32768.0, sr, 16000).mean(dim=0)
    err, err2 = float((got.cpu() - want).abs().max()), float((got.cpu() - ref).abs().max())
    print("resample", sr, channels, tuple(got.shape), err, err2)
    assert got.device.type == "cuda" and got.shape == want.shape
    assert err < 2e-6 and err2 < 2e-6
    fit = fit_audio_unet(got, 16000, 60, 1)             # zero-padding to the 800-sample hop also runs on the device tensor
    assert fit.waveform.device.type == "cuda" and fit.waveform.shape[0] % 800 == 0 and fit.window_size == int(got.shape[0] / 16000 * 60)
