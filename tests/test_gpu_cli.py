"""GPU: the inference CLI (reference `script/inference.py` interface) end to end on a synthetic WAV and a synthetic
checkpoint: WAV decode -> fit -> process -> SAID_UNet1D.inference -> CSV, compared with the CPU oracle; and the
`feature_dim > 0` model variant (audio_proj_layer, diffusion.py:106-112, 228-229)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cli_csv_matches_oracle(tmp_path, state_dict):
    import pandas as pd
    from scipy.io import wavfile

    from oracle import said_oracle as O
    from said_b200.synth import synthetic_waveform
    from said_b200.util.blendshape import DEFAULT_BLENDSHAPE_CLASSES

    wav = tmp_path / "clip.wav"
    x = synthetic_waveform(0, 1.03)                      # 16480 samples: not a multiple of 800 -> fit_audio_unet pads
    wavfile.write(str(wav), 16000, np.round(x * 32767).astype(np.int16))
    ckpt = tmp_path / "weights.pth"
    torch.save(state_dict, str(ckpt))
    out = tmp_path / "out.csv"
    env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, "compat") + os.pathsep + ROOT)
    seed_wrapper = (
        "import sys, runpy, torch; torch.manual_seed(0); sys.argv = sys.argv[1:]; "
        "runpy.run_path(sys.argv[0], run_name='__main__')"
    )
    cmd = [sys.executable, "-c", seed_wrapper, os.path.join(ROOT, "script", "infer_cli.py"), "--weights_path", str(ckpt),
           "--audio_path", str(wav), "--output_path", str(out), "--num_steps", "10", "--device", "cuda:0"]
    subprocess.run(cmd, check=True, env=env, cwd=str(tmp_path), timeout=600)
    df = pd.read_csv(str(out))
    assert list(df.columns) == DEFAULT_BLENDSHAPE_CLASSES
    got = torch.from_numpy(df.values.astype(np.float32))
    # oracle on the same decoded / padded / normalised waveform and the same CUDA noise draw
    from said_b200.util.audio import fit_audio_unet, load_audio

    w = load_audio(str(wav), 16000)
    fit = fit_audio_unet(w, 16000, 60, 1)
    assert got.shape == (fit.window_size, 32) and fit.window_size == 61
    wp = O.process_audio(fit.waveform.numpy())
    T = int(wp.shape[1] / 16000 * 60)
    torch.manual_seed(0)
    noise = torch.randn(1, T, 32, device="cuda:0").cpu()
    with torch.no_grad():
        ref, _ = O.inference(state_dict, wp, num_inference_steps=10, guidance_scale=2.0, noise=noise)
    err = float((got - ref[0, : fit.window_size]).abs().max())
    print("cli csv vs oracle", err)
    assert err < 5e-4


def test_feature_dim_variant_vs_oracle():
    from oracle import said_oracle as O
    from said_b200.model.diffusion import SAID_UNet1D
    from said_b200.synth import synthetic_batch, synthetic_state_dict

    sd = synthetic_state_dict(seed=3, feature_dim=256)
    m = SAID_UNet1D(feature_dim=256)
    m.load_state_dict(sd)
    m.to("cuda:0").eval()
    wave = synthetic_batch(2, 1.0)
    g = torch.Generator().manual_seed(4)
    noise = torch.randn(2, 60, 32, generator=g)
    with torch.no_grad():
        out = m._run(wave.to("cuda:0"), noise.to("cuda:0"), None, None, 10, 1.0, 2.0, 0.0, 0.0, 60, False, False, None).result.cpu()
        ref, _ = O.inference(sd, wave, num_inference_steps=10, guidance_scale=2.0, noise=noise)
    err = float((out - ref).abs().max())
    print("feature_dim=256 vs oracle", err)
    assert err < 5e-4
