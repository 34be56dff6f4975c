"""CPU: host-side logic of the drop-in boundary (no CUDA)."""
import ast
import json
import os
import sys

import numpy as np
import pytest
import torch

from said_b200 import scheduler as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_step_table_matches_scheduler_step(golden_dir):
    """Rows of the table drive the CUDA kernel; evaluating the same formulas on the CPU with a row must
    reproduce DDIMScheduler.step bit for bit."""
    kat = np.load(os.path.join(golden_dir, "scheduler_kat.npz"))
    x, e, z = (torch.from_numpy(kat[k]) for k in ("step_x", "step_e", "step_z"))
    for pt in ("epsilon", "sample", "v_prediction"):
        sch = S.DDIMScheduler(1000, beta_schedule="squaredcos_cap_v2", prediction_type=pt)
        sch.set_timesteps(50)
        for t in (980, 500, 0):
            for eta in (0.0, 0.5):
                sa, sb, sap, dirc, sigma, clip, bsa, bsb = (torch.tensor(v) for v in S.ddim_step_table(sch, [t], eta)[0])
                if pt == "epsilon":
                    x0, eps = (x - sb * e) / sa, e
                elif pt == "sample":
                    x0, eps = e, (x - sa * e) / sb
                else:
                    x0, eps = sa * x - sb * e, sa * e + sb * x
                prev = sap * x0.clamp(-clip, clip) + dirc * eps
                if eta > 0:
                    prev = prev + sigma * z
                assert np.array_equal(prev.numpy(), kat[f"step_{pt}_{t}_{eta}"]), (pt, t, eta)
                assert (float(bsa), float(bsb)) == (1.0, 0.0)


def test_timesteps_and_blend_columns(golden_dir):
    kat = np.load(os.path.join(golden_dir, "scheduler_kat.npz"))
    sch = S.DDIMScheduler(1000, beta_schedule="squaredcos_cap_v2")
    for n in (10, 50, 100, 1000):
        sch.set_timesteps(n)
        assert np.array_equal(sch.timesteps.numpy(), kat[f"timesteps_{n}"])
        assert np.array_equal(sch._timesteps_host, kat[f"timesteps_{n}"])
    sch.set_timesteps(50)
    ts = [int(t) for t in sch.timesteps]
    nxt = ts[1:] + [None]
    tab = S.ddim_step_table(sch, ts, 0.0, nxt)
    assert tuple(tab[-1, 6:8]) == (1.0, 0.0)
    assert tuple(tab[0, 6:8]) == tuple(np.float32(v) for v in S.noise_coefs(sch, ts[1]))
    with pytest.raises(ValueError):
        sch.set_timesteps(2000)


def test_audio_processor_matches_hf_normalisation():
    from said_b200.model.diffusion import AudioProcessor
    from said_b200.synth import normalise_waveform, synthetic_waveform

    w = synthetic_waveform(3, 1.0)
    p = AudioProcessor()
    for inp in (w, torch.from_numpy(w), [w, w]):
        out = p(inp, sampling_rate=16000, return_tensors="pt")["input_values"]
        assert out.dtype == torch.float32 and out.shape[-1] == 16000
        assert np.array_equal(out[0].numpy(), normalise_waveform(w))
    with pytest.raises(ValueError):
        p(w, sampling_rate=8000)


def test_fit_audio_unet_and_csv_roundtrip(tmp_path):
    from said_b200.util.audio import fit_audio_unet, load_audio
    from said_b200.util.blendshape import DEFAULT_BLENDSHAPE_CLASSES, load_blendshape_coeffs, save_blendshape_coeffs

    w = torch.arange(16001, dtype=torch.float32)
    fit = fit_audio_unet(w, 16000, 60, 1)
    assert fit.window_size == 60 and fit.waveform.shape[0] == 16800 and float(fit.waveform[16001:].abs().sum()) == 0.0
    fit = fit_audio_unet(w[:16000], 16000, 60, 1)
    assert fit.window_size == 60 and fit.waveform.shape[0] == 16000
    coeffs = np.random.default_rng(0).random((7, 32)).astype(np.float32)
    path = str(tmp_path / "c.csv")
    save_blendshape_coeffs(coeffs, DEFAULT_BLENDSHAPE_CLASSES, path)
    with open(path) as f:
        assert f.readline().strip().split(",") == DEFAULT_BLENDSHAPE_CLASSES
    assert torch.allclose(load_blendshape_coeffs(path), torch.from_numpy(coeffs), atol=1e-7)
    # WAV loader fallback (int16 PCM)
    from scipy.io import wavfile

    wav = str(tmp_path / "a.wav")
    wavfile.write(wav, 16000, (np.sin(np.arange(1600) / 10.0) * 20000).astype(np.int16))
    x = load_audio(wav, 16000)
    assert x.shape == (1600,) and x.dtype == torch.float32 and float(x.abs().max()) <= 1.0


def test_module_shell_state_dict_layout(golden_dir, state_dict):
    from said_b200.model.diffusion import SAID_UNet1D

    m = SAID_UNet1D()
    with open(os.path.join(golden_dir, "state_dict_layout.json")) as f:
        layout = json.load(f)
    assert {k: list(v.shape) for k, v in m.state_dict().items()} == layout    # the reference's 372 keys / shapes
    res = m.load_state_dict(state_dict)
    assert not res.missing_keys and not res.unexpected_keys
    # new-style weight-norm keys are accepted
    sd2 = dict(state_dict)
    sd2["audio_encoder.encoder.pos_conv_embed.conv.parametrizations.weight.original0"] = sd2.pop(
        "audio_encoder.encoder.pos_conv_embed.conv.weight_g")
    sd2["audio_encoder.encoder.pos_conv_embed.conv.parametrizations.weight.original1"] = sd2.pop(
        "audio_encoder.encoder.pos_conv_embed.conv.weight_v")
    m.load_state_dict(sd2)
    assert m.sampling_rate == 16000 and m.denoiser.in_channels == 32 and m.null_cond_emb.shape == (1, 1, 768)
    assert m.noise_scheduler.config.num_train_timesteps == 1000 and m.noise_scheduler.init_noise_sigma == 1.0
    # zero_module tensors are zero in a fresh model, like the reference
    fresh = SAID_UNet1D()
    fsd = fresh.state_dict()
    assert float(fsd["denoiser.model.out.2.weight"].abs().sum()) == 0.0
    assert float(fsd["denoiser.model.input_blocks.0.0.weight"].abs().sum()) > 0.0
    with pytest.raises(RuntimeError, match="CUDA"):
        m.inference(torch.zeros(1, 16000), num_inference_steps=2)


def test_feature_dim_variant_layout():
    from said_b200.model.diffusion import SAID_UNet1D

    m = SAID_UNet1D(feature_dim=256)
    sd = m.state_dict()
    assert sd["audio_proj_layer.weight"].shape == (256, 768) and sd["null_cond_emb"].shape == (1, 1, 256)
    assert sd["denoiser.model.input_blocks.1.1.transformer_blocks.0.attn2.to_k.weight"].shape == (192, 256)


def test_unsupported_encoder_config_fails_loudly():
    from types import SimpleNamespace

    from said_b200.model.diffusion import SAID_UNet1D

    with pytest.raises(NotImplementedError):
        SAID_UNet1D(audio_config=SimpleNamespace(add_adapter=True))
    with pytest.raises(NotImplementedError):
        SAID_UNet1D(audio_config=SimpleNamespace(feat_extract_norm="layer", conv_bias=False))


def test_large_family_state_dict_layout(large_family):
    """SAID_UNet1D(audio_config=<wav2vec2-large family>) exposes exactly transformers' parameter names for that
    configuration (conv biases, one LayerNorm per conv layer) and loads them strictly."""
    from said_b200.model.diffusion import SAID_UNet1D

    cfg, sd = large_family
    m = SAID_UNet1D(audio_config=cfg)
    own = {k for k in m.state_dict() if k.startswith("audio_encoder.")}
    alias = {"audio_encoder.encoder.pos_conv_embed.conv.parametrizations.weight.original0": "audio_encoder.encoder.pos_conv_embed.conv.weight_g",
             "audio_encoder.encoder.pos_conv_embed.conv.parametrizations.weight.original1": "audio_encoder.encoder.pos_conv_embed.conv.weight_v"}
    theirs = {alias.get(k, k) for k in sd}
    assert own == theirs, (sorted(own - theirs)[:5], sorted(theirs - own)[:5])
    assert "audio_encoder.feature_extractor.conv_layers.3.layer_norm.weight" in own
    assert "audio_encoder.feature_extractor.conv_layers.0.conv.bias" in own
    full = m.state_dict()
    full.update({alias.get(k, k): v for k, v in sd.items()})
    m.load_state_dict(full, strict=True)
    assert m._audio_dims.stable_layer_norm and m.denoiser.model if hasattr(m.denoiser, "model") else True


def test_compat_imports_resolve():
    sys.path.insert(0, os.path.join(ROOT, "compat"))
    try:
        for mod in [m for m in list(sys.modules) if m == "said" or m.startswith("said.") or m.startswith("dataset")]:
            del sys.modules[mod]
        from dataset.dataset_voca import BlendVOCADataset
        from said.model.diffusion import SAID_UNet1D  # noqa: F401
        from said.util.audio import fit_audio_unet, load_audio  # noqa: F401
        from said.util.blendshape import load_blendshape_coeffs, save_blendshape_coeffs, save_blendshape_coeffs_image  # noqa: F401

        assert len(BlendVOCADataset.default_blendshape_classes) == 32
        ref = "/root/reference/script/inference.py"
        if os.path.exists(ref):
            # every `from said... / dataset... / diffusers import X` of the reference's script resolves under compat/
            tree = ast.parse(open(ref).read())
            import importlib

            for node in ast.walk(tree):
                if isinstance(node, ast.ImportFrom) and node.module.split(".")[0] in ("said", "dataset", "diffusers"):
                    if node.module.startswith("diffusers"):
                        try:
                            import diffusers  # noqa: F401
                        except ImportError:
                            pass
                    mod = importlib.import_module(node.module)
                    for alias in node.names:
                        assert hasattr(mod, alias.name), (node.module, alias.name)
            names = [ln.strip() for ln in open("/root/reference/data/ARKit_blendshapes.txt") if ln.strip()]
            assert names == BlendVOCADataset.default_blendshape_classes
    finally:
        sys.path.remove(os.path.join(ROOT, "compat"))


def test_ddpm_scheduler_step_table_and_oracle():
    """DDPMScheduler: the restated step, the oracle's restatement and the float32 coefficient table
    handed to the CUDA kernel are the same arithmetic bit for bit; analytic anchors at the last step (t = 0:
    a_prev = 1, so c0 = 1, c1 = 0, no noise: prev == clipped x0)."""
    from oracle import said_oracle as O

    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 7, 32, generator=g)
    e = torch.randn(2, 7, 32, generator=g)
    z = torch.randn(2, 7, 32, generator=g)
    for pt in ("epsilon", "sample", "v_prediction"):
        sch = S.DDPMScheduler(1000, beta_schedule="squaredcos_cap_v2", prediction_type=pt)
        assert "eta" not in __import__("inspect").signature(sch.step).parameters     # diffusion.py:404-405 depends on this
        sch.set_timesteps(50)
        assert [int(t) for t in sch.timesteps[:3]] == [980, 960, 940] and int(sch.timesteps[-1]) == 0
        for t in (980, 500, 20, 0):
            want = sch.step(e, t, x, variance_noise=z).prev_sample
            ora = O.ddpm_step(e, t, x, sch.alphas_cumprod, 50, pt, variance_noise=z)
            assert torch.equal(want, ora), (pt, t)
            row = torch.from_numpy(S.ddpm_step_table(sch, [t])[0])
            sa, sb, c0, c1, std, clip = row[0], row[1], row[2], row[3], row[4], row[5]
            if pt == "epsilon":
                x0 = (x - sb * e) / sa
            elif pt == "sample":
                x0 = e
            else:
                x0 = sa * x - sb * e
            x0 = x0.clamp(-clip, clip)
            got = c0 * x0 + c1 * x
            if t > 0:
                got = got + std * z
            assert torch.equal(got, want), (pt, t, float((got - want).abs().max()))
            if t == 0:
                assert float(c0) == 1.0 and float(c1) == 0.0 and float(std) == 0.0
                assert torch.equal(want, x0)
    # variance anchor: fixed_small = (1 - a_prev) / (1 - a) * (1 - a / a_prev), computed independently in float64
    sch = S.DDPMScheduler(1000, beta_schedule="squaredcos_cap_v2")
    sch.set_timesteps(1000)
    ac = sch.alphas_cumprod.double()
    t = 500
    v64 = (1 - ac[t - 1]) / (1 - ac[t]) * (1 - ac[t] / ac[t - 1])
    assert abs(float(S.ddpm_step_table(sch, [t])[0][4]) - float(v64**0.5)) < 1e-6


def test_batched_csv_is_byte_identical_to_reference_writer(tmp_path):
    """save_blendshape_coeffs_batch (8(f) rank 1) writes, per clip, exactly the bytes of the reference's pandas writer."""
    from said_b200.util.blendshape import DEFAULT_BLENDSHAPE_CLASSES as C
    from said_b200.util.blendshape import save_blendshape_coeffs, save_blendshape_coeffs_batch

    rng = np.random.default_rng(0)
    x = rng.random((3, 50, 32)).astype(np.float32)
    x[0, 0, :5] = [0.0, 1.0, 1e-5, 0.5, 1e-7]
    x[1, 3, 2] = np.float32(1 / 3)
    x[2] = np.clip(x[2] * 3 - 1, 0, 1)               # plenty of exact 0 / 1 entries, like a clamped result
    save_blendshape_coeffs_batch(x, C, [str(tmp_path / f"b{i}.csv") for i in range(3)])
    for i in range(3):
        save_blendshape_coeffs(x[i], C, str(tmp_path / f"r{i}.csv"))
        assert (tmp_path / f"b{i}.csv").read_bytes() == (tmp_path / f"r{i}.csv").read_bytes()
    with pytest.raises(ValueError):
        save_blendshape_coeffs_batch(x, C, ["only-one.csv"])


def _gemm_tc_items(M, N, G, BN=192, BM=128):
    """Python replica of the work decomposition of tc::gemm_tc_kernel (said_b200/csrc/gemm_tc.cuh, "Tail balancing"):
    returns, per CTA, its list of (row tile, column tile, first column, width)."""
    n_tiles = (N + BN - 1) // BN
    total = ((M + BM - 1) // BM) * n_tiles
    G = min(G, total)
    tq, tr = divmod(total, G)
    S = 1
    if BN == 192 and tq >= 1 and tr > 0:
        for c in (12, 6, 4, 3, 2):
            if c * tr <= G:
                S = c
                break
    out = []
    for cta in range(G):
        items = []
        if S > 1:
            for i in range(tq):
                t = cta * tq + i
                items.append((t // n_tiles, t % n_tiles, 0, BN))
            if cta < tr * S:
                t = G * tq + cta // S
                items.append((t // n_tiles, t % n_tiles, (cta % S) * (BN // S), BN // S))
        else:
            n = tq + (1 if cta < tr else 0)
            t0 = cta * tq + min(cta, tr)
            for i in range(n):
                t = t0 + i
                items.append((t // n_tiles, t % n_tiles, 0, BN))
        out.append(items)
    return out


@pytest.mark.parametrize("M,N", [(38400, 192), (19200, 192), (38400, 576), (38400, 1536), (19200, 576), (37888, 192),
                                 (128 * 149, 192), (128 * 221, 192), (128 * 295, 1536), (4800, 192), (600, 576)])
def test_gemm_tile_and_sliver_decomposition_covers_output_once(M, N):
    """Every (row tile, 16-column group) of the output is produced by exactly one work item, slivers are multiples of 16
    columns, and no CTA gets more than ceil(tiles / CTAs) items (the point of cutting the leftover tiles)."""
    G = 148
    per_cta = _gemm_tc_items(M, N, G)
    seen = {}
    for items in per_cta:
        for mt, nt, n0, nw in items:
            assert nw % 16 == 0 and 16 <= nw <= 192 and n0 % 16 == 0 and n0 + nw <= 192
            for c in range(n0 // 16, (n0 + nw) // 16):
                key = (mt, nt, c)
                assert key not in seen
                seen[key] = True
    rows, cols = (M + 127) // 128, (N + 191) // 192
    assert len(seen) == rows * cols * 12
    total = rows * cols
    assert max(len(i) for i in per_cta) <= -(-total // min(G, total))


def test_compat_diffusers_defers_to_an_installed_package(tmp_path):
    """compat/ first on the path must not shadow a real `diffusers` further down: the stand-in hands the import over."""
    import subprocess

    fake = tmp_path / "site" / "diffusers"
    fake.mkdir(parents=True)
    (fake / "__init__.py").write_text("MARK = 'real'\nclass DDIMScheduler:\n    origin = 'real'\n")
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "from diffusers import DDIMScheduler; import diffusers\n"
            "print(DDIMScheduler.origin, diffusers.MARK)") % (ROOT, str(tmp_path / "site"), os.path.join(ROOT, "compat"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    assert out.stdout.split() == ["real", "real"]
    # without another diffusers on the path the restated scheduler is exported
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "from diffusers import DDIMScheduler; print(DDIMScheduler.__module__)") % (ROOT, os.path.join(ROOT, "compat"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    assert out.stdout.strip() == "said_b200.scheduler"


def test_audio_config_defaults_and_eps_check():
    from said_b200.model.params import Wav2Vec2Dims
    import types

    d = Wav2Vec2Dims(None)
    cfg = d.as_config()
    assert cfg.hidden_size == 768 and cfg.num_hidden_layers == 12 and cfg.conv_kernel[0] == 10
    d.check_supported()
    bad = Wav2Vec2Dims(types.SimpleNamespace(layer_norm_eps=1e-6))
    with pytest.raises(NotImplementedError):
        bad.check_supported()
