"""CPU: the oracle restatement against the golden vectors the UNMODIFIED reference produced
(tests/golden/make_golden.py) -- this is what pins the oracle."""
import json
import os

import numpy as np
import torch

from oracle import said_oracle as O


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def maxdiff(a, b):
    return float((torch.as_tensor(a).double() - torch.as_tensor(b).double()).abs().max())


def test_alpha_table_anchors():
    """SURVEY.md Appendix B anchors of the squaredcos_cap_v2 table (float32)."""
    ac = O.ddim_alphas_cumprod(1000)
    anchors = {0: 0.999958694, 100: 0.971575618, 500: 0.492285043, 900: 2.361610718e-2, 980: 8.765292e-4,
               990: 1.967173e-4, 998: 2.428766e-6, 999: 2.428735e-9}
    for t, v in anchors.items():
        assert abs(float(ac[t]) - v) <= 2e-7 * max(v, 1e-9) + 1e-12, (t, float(ac[t]), v)


def test_timestep_grids():
    for n, first in ((10, 900), (50, 980), (100, 990), (1000, 999)):
        ts = O.ddim_timesteps(n)
        assert ts[0] == first and ts[-1] == 0 and len(ts) == n


def test_scheduler_kat(golden_dir):
    kat = load(golden_dir, "scheduler_kat.npz")
    assert np.array_equal(kat["alphas_cumprod"], O.ddim_alphas_cumprod(1000).numpy())
    x, e, z = (torch.from_numpy(kat[k]) for k in ("step_x", "step_e", "step_z"))
    for pt in ("epsilon", "sample", "v_prediction"):
        for t in (980, 500, 0):
            for eta in (0.0, 0.5):
                out = O.ddim_step(e, t, x, O.ddim_alphas_cumprod(), 50, pt, eta, variance_noise=z)
                assert np.array_equal(out.numpy(), kat[f"step_{pt}_{t}_{eta}"])
    assert np.array_equal(O.ddim_add_noise(x, e, [580, 580], O.ddim_alphas_cumprod()).numpy(), kat["add_noise_580"])
    assert np.array_equal(O.rescale_noise_cfg(x, e, 0.7).numpy(), kat["rescale_0.7"])


def test_denoiser_forward_vs_reference(golden_dir, state_dict):
    gd = load(golden_dir, "denoiser_forward.npz")
    taps = {}
    with torch.no_grad():
        y = O.denoiser_forward(state_dict, torch.from_numpy(gd["x"]), torch.from_numpy(gd["t"]), torch.from_numpy(gd["ctx"]), taps=taps)
    assert maxdiff(y, gd["y"]) < 2e-6     # 0.0 on the machine that made the goldens; thread count may differ
    for name in ("input_blocks.0", "input_blocks.1.0", "input_blocks.1.1", "middle_block.2", "output_blocks.1.1"):
        assert maxdiff(taps[name], gd["act_" + name]) < 2e-5


def test_alignment_mask_is_three_wide():
    for T in (60, 300):
        m = O.alignment_mask(1, T, T)[0]
        for i in (0, 1, T // 2, T - 1):
            keep = (~m[i]).nonzero().flatten().tolist()
            assert keep == list(range(max(i - 1, 0), min(i + 2, T)))


def test_audio_encoder_vs_reference(golden_dir, state_dict):
    gd = load(golden_dir, "audio_encoder_1s.npz")
    taps = {}
    with torch.no_grad():
        emb = O.wav2vec2_forward(state_dict, torch.from_numpy(gd["wave"]), 60, taps=taps)
    assert maxdiff(taps["conv6"], gd["conv_feats"]) < 1e-5
    assert maxdiff(emb, gd["emb"]) < 2e-5


def test_config1_chain_vs_reference(golden_dir, state_dict):
    """BASELINE config 1 (1 s, 10 DDIM steps, epsilon, CFG 2.0) with the reference's audio embedding."""
    gd = load(golden_dir, "config1_1s_10steps_eps.npz")
    with torch.no_grad():
        res, inter = O.inference(state_dict, torch.from_numpy(gd["wave"]), num_inference_steps=10, guidance_scale=2.0,
                                 save_intermediate=True, noise=torch.from_numpy(gd["noise"]), audio_emb=torch.from_numpy(gd["emb"]))
    assert maxdiff(res, gd["result"]) < 2e-4
    assert max(maxdiff(a, b) for a, b in zip(inter, gd["intermediates"])) < 2e-4


def test_editing_keeps_masked_region(golden_dir, state_dict):
    gd = load(golden_dir, "editing_1s_50steps.npz")
    w = load(golden_dir, "config1_1s_10steps_eps.npz")
    init, mask = torch.from_numpy(gd["init"]), torch.from_numpy(gd["mask_shape_s0.6"])
    with torch.no_grad():
        res, _ = O.inference(state_dict, torch.from_numpy(w["wave"]), init_samples=init, mask=mask, num_inference_steps=50,
                             strength=0.1, guidance_scale=2.0, noise=torch.from_numpy(gd["noise_shape_s0.6"]),
                             audio_emb=torch.from_numpy(w["emb"]))
    assert torch.equal(res[mask.bool()], init.clamp(0, 1)[mask.bool()])


def test_floors_recorded(golden_dir):
    with open(os.path.join(golden_dir, "floors.json")) as f:
        d = json.load(f)
    assert d["report"]["denoiser_oracle_vs_ref"] == 0.0
    assert d["report"]["cfg1_loop_oracle_vs_ref_same_emb"] == 0.0
    assert d["report"]["chain1000_v_prediction_oracle_vs_ref"] == 0.0


def test_oracle_large_family_encoder_matches_reference(golden_dir, large_family):
    """wav2vec2-large family (layer-norm feature extractor, conv bias, stable layer norm): the oracle
    against the output of the reference's own ModifiedWav2Vec2Model (tests/golden/make_golden_large.py)."""
    import numpy as np
    import torch

    from oracle import said_oracle as O

    _, sd = large_family
    gd = np.load(os.path.join(golden_dir, "audio_encoder_large_family_1s.npz"))
    with torch.no_grad():
        emb = O.wav2vec2_forward(sd, torch.from_numpy(gd["wave"]), 60, stable_layer_norm=True)
    err = float((emb - torch.from_numpy(gd["emb"])).abs().max())
    assert err < 1e-4, err
