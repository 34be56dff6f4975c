"""Generate the golden vectors under tests/golden/ from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py [--skip-long]

What runs here is the reference's own code: ``said/model/{diffusion,unet_1d_condition,wav2vec2}.py`` and
``said/model/ldm/*`` imported from /root/reference through a synthetic package (its ``said/__init__.py``
drags in librosa/cvxopt/trimesh, which are not installed), with transformers 5.5.0 supplying
``Wav2Vec2Model``.  ``diffusers`` is not installed, so ``said/model/diffusion.py`` gets a stand-in module
whose ``DDIMScheduler`` is ``said_b200.scheduler.DDIMScheduler`` (the restated v0.19 algorithm):
the scheduler boundary is therefore PARITY UNPINNED (see oracle/said_oracle.py), everything else is the
reference verbatim.

The script (1) writes the reference's outputs as .npz fixtures, (2) checks the oracle restatement
against them and records the measured deviations and fp32-vs-fp64 floors in ``floors.json``.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import sys
import time
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from oracle import said_oracle as O  # noqa: E402
from said_b200 import scheduler as S  # noqa: E402
from said_b200.synth import normalise_waveform, synthetic_state_dict, synthetic_waveform  # noqa: E402


def load_reference():
    """Import the reference's model package without executing ``said/__init__.py``."""
    d = types.ModuleType("diffusers")
    d.DDIMScheduler = S.DDIMScheduler
    d.SchedulerMixin = S.SchedulerMixin
    sys.modules["diffusers"] = d
    for name in ("diffusers.pipelines", "diffusers.pipelines.stable_diffusion",
                 "diffusers.pipelines.stable_diffusion.pipeline_stable_diffusion"):
        sys.modules[name] = types.ModuleType(name)
    sys.modules["diffusers.pipelines.stable_diffusion.pipeline_stable_diffusion"].rescale_noise_cfg = O.rescale_noise_cfg
    said = types.ModuleType("said")
    said.__path__ = [REF + "/said"]
    sys.modules["said"] = said
    model = types.ModuleType("said.model")
    model.__path__ = [REF + "/said/model"]
    sys.modules["said.model"] = model
    return importlib.import_module("said.model.diffusion")


class OfflineProcessor:
    """Stands in for ``Wav2Vec2Processor.from_pretrained`` (needs the network): the real HF feature
    extractor, constructed locally with the wav2vec2-base-960h preprocessor settings."""

    def __init__(self):
        from transformers import Wav2Vec2FeatureExtractor

        self.feature_extractor = Wav2Vec2FeatureExtractor(
            feature_size=1, sampling_rate=16000, padding_value=0.0, do_normalize=True, return_attention_mask=False
        )

    def __call__(self, *a, **k):
        return self.feature_extractor(*a, **k)


class RandnRecorder:
    """Wraps torch.randn to log every draw the reference makes inside inference()."""

    def __init__(self):
        self.draws = []
        self._orig = torch.randn

    def __enter__(self):
        def rec(*a, **k):
            out = self._orig(*a, **k)
            self.draws.append(out.detach().clone())
            return out

        torch.randn = rec
        return self

    def __exit__(self, *exc):
        torch.randn = self._orig


def build_model(ref, sd, prediction_type):
    m = ref.SAID_UNet1D(audio_processor=OfflineProcessor(), prediction_type=prediction_type)
    missing = m.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return m.eval()


def maxdiff(a, b):
    return float((torch.as_tensor(a).double() - torch.as_tensor(b).double()).abs().max())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--skip-long", action="store_true", help="skip the 1000-step 5 s chains")
    args = ap.parse_args()
    torch.set_num_threads(os.cpu_count() or 1)
    ref = load_reference()
    sd = synthetic_state_dict(seed=0)
    sd64 = {k: v.double() for k, v in sd.items()}
    floors = {}
    report = {}

    # ---------------------------------------------------------------- state-dict layout
    model = build_model(ref, sd, "epsilon")
    ref_sd = model.state_dict()
    want = {k: tuple(v.shape) for k, v in sd.items()}
    alias = {"weight_g": "parametrizations.weight.original0", "weight_v": "parametrizations.weight.original1"}
    got = {}
    for k, v in ref_sd.items():
        for old, new in alias.items():
            k = k.replace(new, old)
        got[k] = tuple(v.shape)
    assert got == want, set(got.items()) ^ set(want.items())
    with open(os.path.join(HERE, "state_dict_layout.json"), "w") as f:
        json.dump({k: list(s) for k, s in want.items()}, f, indent=0, sort_keys=True)
    ck = {k: float(sd[k].double().sum()) for k in ("null_cond_emb", "denoiser.model.out.2.weight",
                                                    "audio_encoder.encoder.layers.11.final_layer_norm.bias")}
    report["weights_checksum"] = ck

    # ---------------------------------------------------------------- scheduler KATs
    sch = S.DDIMScheduler(num_train_timesteps=1000, beta_schedule="squaredcos_cap_v2")
    kat = {"alphas_cumprod": sch.alphas_cumprod.numpy()}
    for n in (10, 50, 100, 1000):
        sch.set_timesteps(n)
        kat[f"timesteps_{n}"] = sch.timesteps.numpy()
        assert np.array_equal(kat[f"timesteps_{n}"], O.ddim_timesteps(n))
    assert torch.equal(sch.alphas_cumprod, O.ddim_alphas_cumprod(1000))
    g = torch.Generator().manual_seed(11)
    xs = torch.randn(2, 7, 32, generator=g)
    es = torch.randn(2, 7, 32, generator=g)
    zs = torch.randn(2, 7, 32, generator=g)
    kat["step_x"], kat["step_e"], kat["step_z"] = xs.numpy(), es.numpy(), zs.numpy()
    for pt in ("epsilon", "sample", "v_prediction"):
        s2 = S.DDIMScheduler(num_train_timesteps=1000, beta_schedule="squaredcos_cap_v2", prediction_type=pt)
        s2.set_timesteps(50)
        for t in (980, 500, 0):
            for eta in (0.0, 0.5):
                out = s2.step(es, t, xs, eta=eta, variance_noise=zs).prev_sample
                o2 = O.ddim_step(es, t, xs, O.ddim_alphas_cumprod(), 50, pt, eta, variance_noise=zs)
                assert torch.equal(out, o2), (pt, t, eta)
                kat[f"step_{pt}_{t}_{eta}"] = out.numpy()
    kat["add_noise_580"] = sch.add_noise(xs, es, torch.tensor([580, 580])).numpy()
    kat["rescale_0.7"] = O.rescale_noise_cfg(xs, es, 0.7).numpy()
    np.savez_compressed(os.path.join(HERE, "scheduler_kat.npz"), **kat)

    # ---------------------------------------------------------------- denoiser forward (T=60, B=2)
    g = torch.Generator().manual_seed(21)
    x = torch.randn(2, 60, 32, generator=g)
    ctx = torch.randn(2, 60, 768, generator=g)
    ts = torch.tensor([7, 500])
    acts = {}
    hooks = []
    names = ["input_blocks.0", "input_blocks.1.0", "input_blocks.1.1", "middle_block.0", "middle_block.1",
             "middle_block.2", "output_blocks.0.0", "output_blocks.0.1", "output_blocks.1.0", "output_blocks.1.1"]
    mods = dict(model.denoiser.model.named_modules())
    for n in names:
        hooks.append(mods[n].register_forward_hook(lambda m, i, o, n=n: acts.__setitem__(n, o.detach().clone())))
    with torch.no_grad():
        y = model.denoiser(x, ts, ctx)
    for h in hooks:
        h.remove()
    taps = {}
    with torch.no_grad():
        yo = O.denoiser_forward(sd, x, ts, ctx, taps=taps)
        yo64 = O.denoiser_forward(sd64, x.double(), ts, ctx.double())
    report["denoiser_oracle_vs_ref"] = maxdiff(y, yo)
    floors["denoiser_fp32_vs_fp64"] = maxdiff(y, yo64)
    for n in names:
        report[f"denoiser_tap_{n}"] = maxdiff(acts[n], taps[n])
    assert report["denoiser_oracle_vs_ref"] < 2e-5, report
    np.savez_compressed(
        os.path.join(HERE, "denoiser_forward.npz"),
        x=x.numpy(), ctx=ctx.numpy(), t=ts.numpy(), y=y.numpy(), y64=yo64.float().numpy(),
        **{"act_" + n: acts[n].numpy() for n in ("input_blocks.0", "input_blocks.1.0", "input_blocks.1.1",
                                                   "middle_block.2", "output_blocks.1.1")},
    )
    print("denoiser:", report["denoiser_oracle_vs_ref"], floors["denoiser_fp32_vs_fp64"], float(y.abs().mean()))

    # ---------------------------------------------------------------- audio encoder (1 s -> 60 frames)
    wav1 = synthetic_waveform(0, 1.0)
    wp1 = model.process_audio(wav1)
    assert maxdiff(wp1, torch.from_numpy(normalise_waveform(wav1))[None]) == 0.0
    with torch.no_grad():
        emb_ref = model.get_audio_embedding(wp1, 60)
        feats_ref = model.audio_encoder.feature_extractor(wp1)
        taps = {}
        emb_o = O.wav2vec2_forward(sd, wp1, 60, taps=taps)
        emb_o64 = O.wav2vec2_forward(sd64, wp1.double(), 60)
    report["encoder_oracle_vs_ref"] = maxdiff(emb_ref, emb_o)
    report["encoder_conv_oracle_vs_ref"] = maxdiff(feats_ref, taps["conv6"])
    floors["encoder_fp32_vs_fp64"] = maxdiff(emb_ref, emb_o64)
    assert report["encoder_oracle_vs_ref"] < 1e-3, report
    np.savez_compressed(os.path.join(HERE, "audio_encoder_1s.npz"), wave=wp1.numpy(), emb=emb_ref.numpy(),
                        emb64=emb_o64.float().numpy(), conv_feats=feats_ref.numpy())
    print("encoder:", report["encoder_oracle_vs_ref"], floors["encoder_fp32_vs_fp64"], float(emb_ref.abs().mean()))

    # ---------------------------------------------------------------- config 1: 1 s, 10 DDIM steps, eps
    def run_ref(m, wp, seed=0, **kw):
        torch.manual_seed(seed)
        with RandnRecorder() as rec, torch.no_grad():
            out = m.inference(waveform_processed=wp, **kw)
        return out, rec.draws

    out, draws = run_ref(model, wp1, num_inference_steps=10, guidance_scale=2.0, save_intermediate=True)
    noise = draws[0]
    res_o, inter_o = O.inference(sd, wp1, num_inference_steps=10, guidance_scale=2.0, save_intermediate=True, noise=noise)
    res_o64, _ = O.inference(sd64, wp1.double(), num_inference_steps=10, guidance_scale=2.0, noise=noise)
    # same loop fed the reference's own audio embedding: isolates the loop restatement from the
    # encoder's SDPA-vs-explicit-softmax rounding difference
    res_same, inter_same = O.inference(sd, wp1, num_inference_steps=10, guidance_scale=2.0, save_intermediate=True,
                                       noise=noise, audio_emb=emb_ref)
    report["cfg1_loop_oracle_vs_ref_same_emb"] = max(
        maxdiff(out.result, res_same), max(maxdiff(a, b) for a, b in zip(out.intermediates, inter_same)))
    assert report["cfg1_loop_oracle_vs_ref_same_emb"] == 0.0, report
    report["cfg1_oracle_vs_ref"] = maxdiff(out.result, res_o)
    report["cfg1_intermediates_oracle_vs_ref"] = max(maxdiff(a, b) for a, b in zip(out.intermediates, inter_o))
    floors["cfg1_fp32_vs_fp64"] = maxdiff(out.result, res_o64)
    np.savez_compressed(os.path.join(HERE, "config1_1s_10steps_eps.npz"), wave=wp1.numpy(), noise=noise.numpy(),
                        result=out.result.numpy(), intermediates=torch.stack(out.intermediates).numpy(),
                        result64=res_o64.float().numpy(), emb=emb_ref.numpy())
    print("cfg1:", report["cfg1_oracle_vs_ref"], floors["cfg1_fp32_vs_fp64"])

    # single-branch (guidance <= 1), rescale + eta > 0, strength < 1 (generation)
    out, draws = run_ref(model, wp1, num_inference_steps=10, guidance_scale=1.0)
    res_o, _ = O.inference(sd, wp1, num_inference_steps=10, guidance_scale=1.0, noise=draws[0])
    report["nocfg_oracle_vs_ref"] = maxdiff(out.result, res_o)
    np.savez_compressed(os.path.join(HERE, "nocfg_1s_10steps.npz"), noise=draws[0].numpy(), result=out.result.numpy())

    out, draws = run_ref(model, wp1, num_inference_steps=10, guidance_scale=2.0, guidance_rescale=0.7, eta=0.5,
                         save_intermediate=True)
    eta_noise = torch.stack(draws[1:])
    assert eta_noise.shape[0] == 10
    res_o, _ = O.inference(sd, wp1, num_inference_steps=10, guidance_scale=2.0, guidance_rescale=0.7, eta=0.5,
                           noise=draws[0], eta_noise=eta_noise)
    report["eta_rescale_oracle_vs_ref"] = maxdiff(out.result, res_o)
    np.savez_compressed(os.path.join(HERE, "eta_rescale_1s_10steps.npz"), noise=draws[0].numpy(),
                        eta_noise=eta_noise.numpy(), result=out.result.numpy(),
                        intermediates=torch.stack(out.intermediates).numpy())

    # ---------------------------------------------------------------- editing (1 s, 50 steps)
    g = torch.Generator().manual_seed(31)
    init = (0.4 * torch.rand(1, 60, 32, generator=g)).float()
    m_between = torch.zeros(1, 60, 32)
    m_between[:, :20] = 1
    m_between[:, 40:] = 1
    m_shape = torch.zeros(1, 60, 32)
    m_shape[:, :, :16] = 1
    edit = {"init": init.numpy()}
    for tag, msk, strength in (("between_s1.0", m_between, 1.0), ("shape_s0.6", m_shape, 0.6)):
        out, draws = run_ref(model, wp1, init_samples=init, mask=msk, num_inference_steps=50, strength=strength,
                             guidance_scale=2.0)
        res_o, _ = O.inference(sd, wp1, init_samples=init, mask=msk, num_inference_steps=50, strength=strength,
                               guidance_scale=2.0, noise=draws[0])
        report[f"edit_{tag}_oracle_vs_ref"] = maxdiff(out.result, res_o)
        kept = msk.bool()
        assert torch.equal(out.result[kept], init.clamp(0, 1)[kept])
        edit[f"noise_{tag}"] = draws[0].numpy()
        edit[f"mask_{tag}"] = msk.numpy()
        edit[f"result_{tag}"] = out.result.numpy()
    np.savez_compressed(os.path.join(HERE, "editing_1s_50steps.npz"), **edit)
    print("editing:", {k: v for k, v in report.items() if k.startswith("edit")})

    # ---------------------------------------------------------------- long chains (5 s, 1000 steps)
    if not args.skip_long:
        wp5 = model.process_audio(synthetic_waveform(0, 5.0))
        for pt in ("v_prediction", "sample"):
            m = build_model(ref, sd, pt)
            t0 = time.time()
            out, draws = run_ref(m, wp5, num_inference_steps=1000, guidance_scale=2.0)
            wall = time.time() - t0
            noise = draws[0]
            with torch.no_grad():
                emb5 = m.get_audio_embedding(wp5, 300)
            res_o, pre_o = O.inference(sd, wp5, num_inference_steps=1000, guidance_scale=2.0, noise=noise,
                                       prediction_type=pt, audio_emb=emb5, return_preclamp=True)
            res64, pre64 = O.inference(sd64, wp5.double(), num_inference_steps=1000, guidance_scale=2.0, noise=noise,
                                       prediction_type=pt, audio_emb=emb5.double(), return_preclamp=True)
            report[f"chain1000_{pt}_oracle_vs_ref"] = maxdiff(out.result, res_o)
            floors[f"chain1000_{pt}_fp32_vs_fp64"] = maxdiff(out.result, res64)
            report[f"chain1000_{pt}_ref_wall_s_{os.cpu_count()}cores"] = wall
            np.savez_compressed(os.path.join(HERE, f"chain_5s_1000steps_{pt}.npz"), noise=noise.numpy(),
                                result=out.result.numpy(), result64=res64.float().numpy(),
                                preclamp64=pre64[-1].float().numpy())
            print(pt, report[f"chain1000_{pt}_oracle_vs_ref"], floors[f"chain1000_{pt}_fp32_vs_fp64"], wall)

    with open(os.path.join(HERE, "floors.json"), "w") as f:
        json.dump({"floors": floors, "report": report, "torch": torch.__version__,
                   "threads": torch.get_num_threads()}, f, indent=1, sort_keys=True)
    print(json.dumps(report, indent=1))


if __name__ == "__main__":
    main()
