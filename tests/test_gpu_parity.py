"""GPU parity tests: the sm_100a path (through the C ABI) against the golden vectors produced by the
UNMODIFIED reference (tests/golden/make_golden.py) and against the CPU oracle.

Tolerances are absolute, on tensors whose entries are O(1); each is stated next to the reference's own
fp32-vs-fp64 floor for the same quantity (tests/golden/floors.json).
"""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def maxdiff(a, b):
    a = torch.as_tensor(a).detach().cpu().double()
    b = torch.as_tensor(b).detach().cpu().double()
    return float((a - b).abs().max())


def floors(golden_dir):
    with open(os.path.join(golden_dir, "floors.json")) as f:
        return json.load(f)["floors"]


# ------------------------------------------------------------------------------------------------ scheduler
def test_ddim_step_bit_exact(golden_dir, gpu_model):
    """Fused step kernel == restated DDIMScheduler.step, bit for bit, all prediction types, eta 0 / 0.5."""
    from said_b200 import scheduler as S

    kat = load(golden_dir, "scheduler_kat.npz")
    eng = gpu_model()._engine(torch.device(DEV))
    x = torch.from_numpy(kat["step_x"]).to(DEV)
    e = torch.from_numpy(kat["step_e"]).to(DEV)
    z = torch.from_numpy(kat["step_z"]).to(DEV)
    for code, pt in enumerate(("epsilon", "sample", "v_prediction")):
        sch = S.DDIMScheduler(1000, beta_schedule="squaredcos_cap_v2", prediction_type=pt)
        sch.set_timesteps(50)
        for t in (980, 500, 0):
            for eta in (0.0, 0.5):
                row = S.ddim_step_table(sch, [t], eta)[0]
                out = eng.op_ddim_step(e, x, False, 1.0, 0.0, code, row, eta_noise=z if eta > 0 else None)
                want = kat[f"step_{pt}_{t}_{eta}"]
                assert np.array_equal(out.cpu().numpy(), want), (pt, t, eta, maxdiff(out, want))


def test_cfg_rescale_step_matches_oracle(gpu_model):
    from oracle import said_oracle as O
    from said_b200 import scheduler as S

    eng = gpu_model()._engine(torch.device(DEV))
    g = torch.Generator().manual_seed(5)
    B, T, C = 3, 60, 32
    pred = torch.randn(2 * B, T, C, generator=g)
    x = torch.randn(B, T, C, generator=g)
    sch = S.DDIMScheduler(1000, beta_schedule="squaredcos_cap_v2")
    sch.set_timesteps(50)
    row = S.ddim_step_table(sch, [500], 0.0)[0]
    u, c = pred.chunk(2)
    cfg = c + 2.0 * (c - u)
    cfg_r = O.rescale_noise_cfg(cfg, c, 0.7)
    for resc, ref_pred in ((0.0, cfg), (0.7, cfg_r)):
        want = O.ddim_step(ref_pred, 500, x, O.ddim_alphas_cumprod(), 50, "epsilon", 0.0)
        out = eng.op_ddim_step(pred.to(DEV), x.to(DEV), True, 2.0, resc, 0, row)
        tol = 0.0 if resc == 0.0 else 2e-6   # std reduction order differs from torch.std by an ulp or two
        assert maxdiff(out, want) <= tol, (resc, maxdiff(out, want))


# ------------------------------------------------------------------------------------------------ attention
@pytest.mark.parametrize("heads,hd,T,B", [(6, 32, 300, 2), (6, 32, 61, 1), (12, 64, 300, 1), (12, 64, 130, 2), (6, 32, 469, 1)])
def test_self_attention_vs_torch(gpu_model, heads, hd, T, B):
    eng = gpu_model()._engine(torch.device(DEV))
    g = torch.Generator().manual_seed(heads * 1000 + T)
    qkv = torch.randn(B, T, 3 * heads * hd, generator=g).to(DEV)
    out = eng.op_self_attention(qkv, heads, hd)
    q, k, v = qkv.double().chunk(3, dim=-1)
    sh = lambda t: t.reshape(B, T, heads, hd).transpose(1, 2)  # noqa: E731
    ref = torch.softmax(sh(q) @ sh(k).transpose(-1, -2) * hd**-0.5, dim=-1) @ sh(v)
    ref = ref.transpose(1, 2).reshape(B, T, heads * hd)
    assert maxdiff(out, ref) < 3e-6   # fp32 kernel vs fp64 math on O(1) values


@pytest.mark.parametrize("T,B", [(300, 2), (60, 3), (128, 1), (257, 1), (304, 1), (17, 2)])
def test_self_attention_tensor_core_vs_torch(gpu_model, T, B):
    """tcgen05 attention (3xTF32 QK^T and PV, fp32 softmax) against fp64 math; tolerance 2e-5 (fp32 kernel: 3e-6)."""
    eng = gpu_model()._engine(torch.device(DEV))
    heads, hd = 6, 32
    g = torch.Generator().manual_seed(T)
    qkv = torch.randn(B, T, 3 * heads * hd, generator=g).to(DEV)
    out = eng.op_self_attention_tc(qkv, heads)
    q, k, v = qkv.double().chunk(3, dim=-1)
    sh = lambda t: t.reshape(B, T, heads, hd).transpose(1, 2)  # noqa: E731
    ref = torch.softmax(sh(q) @ sh(k).transpose(-1, -2) * hd**-0.5, dim=-1) @ sh(v)
    ref = ref.transpose(1, 2).reshape(B, T, heads * hd)
    err = maxdiff(out, ref)
    print("tc attention", T, err)
    assert err < 2e-5


# ------------------------------------------------------------------------------------------------ denoiser
TAP_ORDER = ["input_blocks.0", "input_blocks.1.0", "input_blocks.1.1", "middle_block.0", "middle_block.1",
             "middle_block.2", "output_blocks.0.0", "output_blocks.0.1", "output_blocks.1.0", "output_blocks.1.1"]


def test_denoiser_forward_golden(golden_dir, gpu_model):
    """One UNet forward vs the reference's own module output and per-block activations.
    Reference fp32-vs-fp64 floor: 2.4e-6; tolerance 2e-5."""
    gd = load(golden_dir, "denoiser_forward.npz")
    eng = gpu_model()._engine(torch.device(DEV))
    x = torch.from_numpy(gd["x"]).to(DEV)
    ctx = torch.from_numpy(gd["ctx"]).to(DEV)
    t = torch.from_numpy(gd["t"])
    out, taps = eng.denoiser_forward(x, t, ctx, taps=True)
    errs = {}
    for i, name in enumerate(TAP_ORDER):
        key = "act_" + name
        if key in gd.files:
            errs[name] = maxdiff(taps[i].transpose(1, 2), gd[key])   # golden activations are (B, C, T)
    errs["out_vs_ref32"] = maxdiff(out, gd["y"])
    errs["out_vs_ref64"] = maxdiff(out, gd["y64"])
    print(errs)
    assert all(v < 2e-5 for v in errs.values()), errs


@pytest.mark.parametrize("mode,tol", [("tf32x3", 2e-4), ("tf32", 3e-2), ("fp32", 2e-5)])
def test_denoiser_forward_precision_modes(golden_dir, gpu_model, mode, tol):
    """Every loader / epilogue of the tcgen05 GEMM (forced on even for this small problem) against the
    reference's activations.  3xTF32 stays within 10x of the fp32 kernel (tensor-core accumulation truncates); single-pass TF32 is
    reported and bounded loosely."""
    gd = load(golden_dir, "denoiser_forward.npz")
    m = gpu_model()
    eng = m._engine(torch.device(DEV))
    eng.set_precision(mode, 1, "fp32")
    try:
        out, taps = eng.denoiser_forward(torch.from_numpy(gd["x"]).to(DEV), torch.from_numpy(gd["t"]),
                                         torch.from_numpy(gd["ctx"]).to(DEV), taps=True)
    finally:
        eng.set_precision(m.precision, -1, m.encoder_precision)
    errs = {name: maxdiff(taps[i].transpose(1, 2), gd["act_" + name]) for i, name in enumerate(TAP_ORDER) if "act_" + name in gd.files}
    errs["out"] = maxdiff(out, gd["y64"])
    print(mode, errs)
    assert all(v < tol for v in errs.values()), (mode, errs)


def test_model_forward_api(golden_dir, gpu_model):
    gd = load(golden_dir, "denoiser_forward.npz")
    m = gpu_model()
    with torch.no_grad():
        y = m(torch.from_numpy(gd["x"]).to(DEV), torch.from_numpy(gd["t"]).to(DEV), torch.from_numpy(gd["ctx"]).to(DEV))
    assert maxdiff(y, gd["y"]) < 2e-5
    # scalar timestep broadcast (diffusion.py:149-152)
    y1 = m(torch.from_numpy(gd["x"]).to(DEV), torch.tensor(500), torch.from_numpy(gd["ctx"]).to(DEV))
    assert maxdiff(y1[1], gd["y"][1]) < 2e-5


# ------------------------------------------------------------------------------------------------ audio encoder
def test_audio_encoder_golden(golden_dir, gpu_model):
    """Wav2Vec2 path vs the reference (transformers) output. Reference fp32-vs-fp64 floor 3.4e-6; tolerance 5e-5."""
    gd = load(golden_dir, "audio_encoder_1s.npz")
    m = gpu_model()
    emb = m.get_audio_embedding(torch.from_numpy(gd["wave"]).to(DEV), 60)
    e32, e64 = maxdiff(emb, gd["emb"]), maxdiff(emb, gd["emb64"])
    print("encoder", e32, e64)
    assert e32 < 5e-5 and e64 < 5e-5


def test_audio_encoder_tensor_core_mode(golden_dir, gpu_model):
    """Opt-in tcgen05 (3xTF32) encoder: 8 x 1 s clips, every GEMM on the tensor cores, against the fp32 mode.
    The 12-layer stack with contractions up to K = 3072 loses about a digit (measured 3e-4 on O(1) features):
    tolerance 1e-3, which is why fp32 is the encoder's default."""
    from said_b200.synth import synthetic_batch

    m = gpu_model()
    wave = synthetic_batch(8, 1.0).to(DEV)
    default = m.encoder_precision
    m.encoder_precision = "fp32"
    ref = m.get_audio_embedding(wave, 60)
    m.encoder_precision, m.tc_min_rows = "tf32x3", 1
    try:
        got = m.get_audio_embedding(wave, 60)
    finally:
        m.encoder_precision, m.tc_min_rows = default, 0
        m._engine(torch.device(DEV)).set_precision(m.precision, -1, default)
    e = maxdiff(got, ref)
    print("encoder tf32x3 vs fp32", e)
    assert 0 < e < 1e-3


@pytest.mark.parametrize("pt", ["v_prediction"])
def test_chain_1000_steps_tensor_core_mode(golden_dir, gpu_model, pt):
    """The 1000-step free-running chain with EVERY denoiser GEMM and attention on the tensor cores (3xTF32),
    forced on for this single clip: tolerance 1e-3 against the reference (the fp32 kernels: 3e-4)."""
    from said_b200.synth import normalise_waveform, synthetic_waveform

    gd = load(golden_dir, f"chain_5s_1000steps_{pt}.npz")
    wave = torch.from_numpy(normalise_waveform(synthetic_waveform(0, 5.0)))[None]
    m = gpu_model(pt)
    default = m.precision
    m.tc_min_rows, m.precision = 1, "tf32x3"
    try:
        out = run(m, wave, gd["noise"], steps=1000)
    finally:
        m.tc_min_rows, m.precision = 0, default
        m._engine(torch.device(DEV)).set_precision(m.precision, -1, m.encoder_precision)
    e32, e64 = maxdiff(out.result, gd["result"]), maxdiff(out.result, gd["result64"])
    print("tensor-core chain", pt, e32, e64)
    assert e32 < 1e-3 and e64 < 1e-3


# ------------------------------------------------------------------------------------------------ chains
def run(m, wave, noise, init=None, mask=None, steps=10, strength=1.0, gs=2.0, gr=0.0, eta=0.0, eta_noise=None,
        inter=False, T=None):
    wave = torch.as_tensor(wave).to(DEV)
    T = T or int(wave.shape[1] / 16000 * 60)
    cv = lambda a: None if a is None else torch.as_tensor(a).to(DEV)  # noqa: E731
    with torch.no_grad():
        return m._run(wave, cv(noise), cv(init), cv(mask), steps, strength, gs, gr, eta, T, inter, False, cv(eta_noise),
                      return_latents=True)


def test_config1_chain_golden(golden_dir, gpu_model):
    """BASELINE config 1: 1 s clip, 10 DDIM steps, epsilon prediction, CFG 2.0, end to end (encoder + loop).
    Reference fp32-vs-fp64 floor on the result: 4.8e-5; tolerance 5e-4 on result and every intermediate."""
    gd = load(golden_dir, "config1_1s_10steps_eps.npz")
    out = run(gpu_model("epsilon"), gd["wave"], gd["noise"], inter=True)
    e_res = maxdiff(out.result, gd["result"])
    e_int = max(maxdiff(a, b) for a, b in zip(out.intermediates, gd["intermediates"]))
    print("config1", e_res, e_int, maxdiff(out.result, gd["result64"]))
    assert len(out.intermediates) == 10
    assert e_res < 5e-4 and e_int < 5e-4


def test_nocfg_and_eta_rescale_golden(golden_dir, gpu_model):
    w = load(golden_dir, "config1_1s_10steps_eps.npz")["wave"]
    m = gpu_model("epsilon")
    g1 = load(golden_dir, "nocfg_1s_10steps.npz")
    out = run(m, w, g1["noise"], gs=1.0)
    assert maxdiff(out.result, g1["result"]) < 5e-4
    g2 = load(golden_dir, "eta_rescale_1s_10steps.npz")
    out = run(m, w, g2["noise"], gr=0.7, eta=0.5, eta_noise=g2["eta_noise"], inter=True)
    e = max(maxdiff(out.result, g2["result"]), max(maxdiff(a, b) for a, b in zip(out.intermediates, g2["intermediates"])))
    print("eta+rescale", e)
    assert e < 1e-3


def test_editing_golden(golden_dir, gpu_model):
    """Editing mode (init_samples + mask), 50 DDIM steps; kept region must equal clamp(init) exactly."""
    w = load(golden_dir, "config1_1s_10steps_eps.npz")["wave"]
    gd = load(golden_dir, "editing_1s_50steps.npz")
    m = gpu_model("epsilon")
    init = torch.from_numpy(gd["init"])
    for tag, strength in (("between_s1.0", 1.0), ("shape_s0.6", 0.6)):
        mask = torch.from_numpy(gd[f"mask_{tag}"])
        out = run(m, w, gd[f"noise_{tag}"], init=init, mask=mask, steps=50, strength=strength)
        res = out.result.cpu()
        kept = mask.bool()
        assert torch.equal(res[kept], init.clamp(0, 1)[kept])
        e = maxdiff(res, gd[f"result_{tag}"])
        print("editing", tag, e)
        assert e < 5e-3   # 50 epsilon-prediction steps with untrained weights amplify rounding (oracle-vs-ref: 1.3e-4)


@pytest.mark.parametrize("pt", ["v_prediction", "sample"])
def test_chain_1000_steps_golden(golden_dir, gpu_model, pt):
    """BASELINE config 2 shape: 5 s clip, 1000 DDIM steps, free-running, against the reference's result.
    Reference fp32-vs-fp64 floors: v_prediction 1.9e-5, sample 2.4e-5; tolerance 3e-4."""
    from said_b200.synth import normalise_waveform, synthetic_waveform

    gd = load(golden_dir, f"chain_5s_1000steps_{pt}.npz")
    wave = torch.from_numpy(normalise_waveform(synthetic_waveform(0, 5.0)))[None]
    out = run(gpu_model(pt), wave, gd["noise"], steps=1000)
    e32, e64 = maxdiff(out.result, gd["result"]), maxdiff(out.result, gd["result64"])
    epre = maxdiff(out.latents, gd["preclamp64"])
    print(pt, e32, e64, epre)
    assert e32 < 3e-4 and e64 < 3e-4 and epre < 3e-4


def test_batched_tensor_core_chain_vs_oracle(gpu_model, state_dict):
    """40 clips x 1 s (4800 denoiser rows: the tcgen05 GEMM path at its production tile shapes), 10 DDIM steps,
    against (a) the fp32 FFMA mode of the same engine and (b) the CPU oracle on two of the clips.
    Tolerance 5e-4 (same as config 1; the reference's own fp32-vs-fp64 floor there is 4.8e-5)."""
    from oracle import said_oracle as O
    from said_b200.synth import synthetic_batch

    m = gpu_model("epsilon")
    wave = synthetic_batch(40, 1.0)
    g = torch.Generator().manual_seed(9)
    noise = torch.randn(40, 60, 32, generator=g)
    res = {}
    default = m.precision
    for mode in ("tf32x3", "fp32", "tf32"):
        m.precision = mode
        try:
            res[mode] = run(m, wave, noise, steps=10).result.cpu()
        finally:
            m.precision = default
    e_modes = maxdiff(res["tf32x3"], res["fp32"])
    e_tf32 = maxdiff(res["tf32"], res["fp32"])
    with torch.no_grad():
        ref, _ = O.inference(state_dict, wave[[0, 39]], num_inference_steps=10, guidance_scale=2.0, noise=noise[[0, 39]])
    e_or = maxdiff(res["tf32x3"][[0, 39]], ref)
    print("batched tc chain: tf32x3 vs fp32", e_modes, "tf32 vs fp32", e_tf32, "tf32x3 vs oracle", e_or)
    assert e_modes < 5e-4 and e_or < 5e-4
    assert e_tf32 < 1e-1


def test_long_clip_mixed_kernels_vs_oracle(gpu_model, state_dict):
    """6 s clips (T = 360 > 304 keys): tcgen05 GEMMs with the FFMA attention kernel; one clip checked against the oracle."""
    from oracle import said_oracle as O
    from said_b200.synth import synthetic_batch

    m = gpu_model("epsilon")
    wave = synthetic_batch(4, 6.0)
    g = torch.Generator().manual_seed(13)
    noise = torch.randn(4, 360, 32, generator=g)
    res = run(m, wave, noise, steps=4).result.cpu()
    with torch.no_grad():
        ref, _ = O.inference(state_dict, wave[[3]], num_inference_steps=4, guidance_scale=2.0, noise=noise[[3]])
    e = maxdiff(res[[3]], ref)
    print("long clip", e)
    assert e < 5e-4


def synthetic_coefficients(batch: int, frames: int, seed: int = 0) -> torch.Tensor:
    """Smooth blendshape-coefficient curves in [0, 0.75] (the range of the reference's data/blendshape_coeffs.zip), used as
    realistic init_samples for the editing path."""
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(frames, dtype=torch.float32)[None, :, None] / 60.0
    f = 0.3 + 2.5 * torch.rand(batch, 1, 32, generator=g)
    ph = 6.2832 * torch.rand(batch, 1, 32, generator=g)
    amp = 0.375 * torch.rand(batch, 1, 32, generator=g)
    return (amp * (1.0 + torch.sin(6.2832 * f * t + ph))).clamp(0.0, 0.75)


def test_bench_shape_chain_vs_oracle(gpu_model, state_dict):
    """The benchmarked configuration itself (BASELINE configs[2] at half batch so the test stays short): 32 clips x 5 s
    (T = 300, 19 200 denoiser rows under CFG with the shared guidance prefix, tail slivers, cluster GroupNorm, tensor-core
    attention), default precision, 4 DDIM steps, against the CPU oracle on the first and the last clip."""
    from oracle import said_oracle as O
    from said_b200.synth import synthetic_batch

    m = gpu_model("epsilon")
    B = 32
    wave = synthetic_batch(B, 5.0)
    g = torch.Generator().manual_seed(17)
    noise = torch.randn(B, 300, 32, generator=g)
    out = run(m, wave, noise, steps=4)
    res, lat = out.result.cpu(), out.latents.cpu()
    sel = [0, B - 1]
    with torch.no_grad():
        ref, pre = O.inference(state_dict, wave[sel], num_inference_steps=4, guidance_scale=2.0, noise=noise[sel], return_preclamp=True)
    e, ep = maxdiff(res[sel], ref), maxdiff(lat[sel], pre[-1])
    print("bench-shape chain vs oracle: result", e, "pre-clamp latents", ep)
    assert bool(torch.isfinite(res).all())
    assert e < 5e-4 and ep < 1e-3


def test_editing_batch_T300_vs_oracle(gpu_model, state_dict):
    """BASELINE configs[4] shape per GPU: editing (init_samples + mask) of 16 clips x 5 s, 50 DDIM steps, both mask styles of
    the reference's demos (in-between frames / half of the blendshapes), strength 1.0 and 0.6, default precision.  Kept
    region bit-equal to clamp(init); the rest against the oracle on two clips (tolerance as test_editing_golden)."""
    from oracle import said_oracle as O
    from said_b200.synth import synthetic_batch

    m = gpu_model("epsilon")
    B, T = 16, 300
    wave = synthetic_batch(B, 5.0)
    init = synthetic_coefficients(B, T, seed=3)
    g = torch.Generator().manual_seed(23)
    noise = torch.randn(B, T, 32, generator=g)
    masks = {"between": torch.zeros(B, T, 32), "shape": torch.zeros(B, T, 32)}
    masks["between"][:, :100] = 1.0
    masks["between"][:, 200:] = 1.0
    masks["shape"][:, :, :16] = 1.0
    sel = [1, B - 1]
    for tag, strength in (("between", 1.0), ("shape", 0.6)):
        mask = masks[tag]
        res = run(m, wave, noise, init=init, mask=mask, steps=50, strength=strength).result.cpu()
        kept = mask.bool()
        assert torch.equal(res[kept], init.clamp(0, 1)[kept]), tag
        with torch.no_grad():
            ref, _ = O.inference(state_dict, wave[sel], init_samples=init[sel], mask=mask[sel], num_inference_steps=50,
                                 strength=strength, guidance_scale=2.0, noise=noise[sel])
        e = maxdiff(res[sel], ref)
        print("editing T=300 batch 16", tag, strength, e)
        assert e < 5e-3


def test_wav2vec2_large_full_size(golden_dir):
    """The REAL wav2vec2-large architecture (hidden 1024, 24 pre-LN layers, 16 heads, ffn 4096, LayerNorm feature extractor with
    conv bias; 315 M parameters) against the output of the reference's own ModifiedWav2Vec2Model
    (tests/golden/make_golden_large.py); weights regenerated from the seed through transformers' Wav2Vec2Model."""
    from transformers import Wav2Vec2Config, Wav2Vec2Model

    from said_b200.model.diffusion import SAID_UNet1D

    gd = load(golden_dir, "audio_encoder_large_full_1s.npz")
    cfg = Wav2Vec2Config(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, intermediate_size=4096,
                         feat_extract_norm="layer", conv_bias=True, do_stable_layer_norm=True,
                         num_conv_pos_embeddings=128, num_conv_pos_embedding_groups=16, vocab_size=32)
    torch.manual_seed(0)
    hf = Wav2Vec2Model(cfg).eval()
    sd = {"audio_encoder." + k: v.detach().clone() for k, v in hf.state_dict().items()}
    got = sum(float(v.double().abs().sum()) for v in sd.values())
    if abs(got - float(gd["weights_abs_sum"])) > 1e-6 * abs(got):
        pytest.skip("transformers initialises Wav2Vec2Model differently here: the golden weights cannot be regenerated")
    del hf
    m = SAID_UNet1D(audio_config=cfg)
    m.load_state_dict(sd, strict=False)
    m.to(DEV).eval()
    with torch.no_grad():
        emb = m.get_audio_embedding(torch.from_numpy(gd["wave"]).to(DEV), 60)
    e32, e64 = maxdiff(emb, gd["emb"]), maxdiff(emb, gd["emb64"])
    print("wav2vec2-large full size", e32, e64, "reference fp32-vs-fp64 floor", float(gd["fp32_vs_fp64"]))
    assert emb.shape == (1, 60, 1024)
    assert e32 < 2e-4 and e64 < 2e-4


def test_step_graph_is_cached(gpu_model):
    """The instantiated per-step CUDA graph is reused by the next denoise() call with the same shapes and scalars, and rebuilt
    when they change; results are unaffected."""
    from said_b200.synth import synthetic_batch

    m = gpu_model("epsilon")
    eng = m._engine(torch.device(DEV))
    wave = synthetic_batch(2, 1.0)
    g = torch.Generator().manual_seed(4)
    noise = torch.randn(2, 60, 32, generator=g)
    a = run(m, wave, noise, steps=6).result.cpu()
    c0 = eng.graph_captures
    b = run(m, wave, noise, steps=6).result.cpu()
    assert eng.graph_captures == c0, "same call: the cached graph must be replayed"
    c = run(m, wave, noise, steps=6, gs=3.0).result.cpu()
    assert eng.graph_captures == c0 + 1, "a different guidance scale is baked into the step kernel: re-capture"
    d = run(m, wave, noise, steps=6).result.cpu()
    assert torch.equal(a, b) and torch.equal(a, d) and not torch.equal(a, c)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_two_engines_in_one_process(gpu_model, state_dict):
    """One process driving two devices (SAID._engines is keyed by device): kernel attributes are configured per device."""
    from said_b200.synth import synthetic_batch

    m = gpu_model("epsilon")
    wave = synthetic_batch(40, 1.0)
    g = torch.Generator().manual_seed(9)
    noise = torch.randn(40, 60, 32, generator=g)
    a = run(m, wave, noise, steps=4).result.cpu()
    m1 = m.to("cuda:1")
    try:
        with torch.no_grad():
            b = m1._run(wave.to("cuda:1"), noise.to("cuda:1"), None, None, 4, 1.0, 2.0, 0.0, 0.0, 60, False, False, None).result.cpu()
    finally:
        m.to(DEV)
    assert torch.equal(a, b)


def test_chunked_variance_noise_loop_is_bit_identical(gpu_model):
    """eta > 0 draws one noise tensor per step (DDIMScheduler.step); the loop then runs in chunks that hold only a bounded number
    of steps' noise (n_loop x B x T x C up front is 2.5 GB at batch 64 x 1000 steps).  Chunking must not change a
    bit: same torch.randn sequence, latents carried between chunks, final un-noised blend only in the last chunk."""
    from said_b200.synth import synthetic_batch

    m = gpu_model("epsilon")
    wave = synthetic_batch(3, 1.0).to(DEV)
    init = synthetic_coefficients(3, 60, seed=8).to(DEV)
    mask = torch.zeros(3, 60, 32, device=DEV)
    mask[:, :20] = 1.0
    outs = []
    saved = m.eta_chunk_bytes
    try:
        for chunk_steps in (1000, 3, 1):
            m.eta_chunk_bytes = chunk_steps * 3 * 60 * 32 * 4
            torch.manual_seed(77)
            with torch.no_grad():
                o = m.inference(wave, init_samples=init, mask=mask, num_inference_steps=10, strength=0.8, guidance_scale=2.0,
                                eta=0.5, save_intermediate=True)
            outs.append((o.result.clone(), torch.stack(o.intermediates).clone()))
    finally:
        m.eta_chunk_bytes = saved
    for r, i in outs[1:]:
        assert torch.equal(r, outs[0][0]) and torch.equal(i, outs[0][1])
    kept = mask.bool()
    assert torch.equal(outs[0][0][kept], init.clamp(0, 1)[kept])


# ------------------------------------------------------------------------------------------------ invariants
def test_invariants(gpu_model):
    from said_b200.synth import synthetic_batch

    m = gpu_model("epsilon")   # (4 clips x 60 frames: every GEMM is below the tensor-core row threshold, one kernel regime)
    wave = synthetic_batch(4, 1.0)
    g = torch.Generator().manual_seed(3)
    noise = torch.randn(4, 60, 32, generator=g)
    full = run(m, wave, noise, steps=10).result
    assert float(full.min()) >= 0.0 and float(full.max()) <= 1.0
    # clips are independent: any sub-batch / permutation reproduces the same rows bit for bit
    perm = torch.tensor([2, 0, 3, 1])
    p = run(m, wave[perm], noise[perm], steps=10).result
    assert torch.equal(p, full[perm])
    halves = torch.cat([run(m, wave[:2], noise[:2], steps=10).result, run(m, wave[2:], noise[2:], steps=10).result])
    assert torch.equal(halves, full)
    # CUDA-graph replay and plain launches are the same program
    m.use_cuda_graph = False
    try:
        plain = run(m, wave, noise, steps=10).result
    finally:
        m.use_cuda_graph = True
    assert torch.equal(plain, full)
    # strength < 1 runs int(N * strength) iterations; guidance <= 1 runs one branch
    out = run(m, wave, noise, steps=10, strength=0.5, inter=True)
    assert len(out.intermediates) == 5
    # empty loop: result = clamp(init)
    out = run(m, wave, noise, steps=10, strength=0.0)
    assert torch.equal(out.result.cpu(), noise.clamp(0, 1))


def test_public_inference_seeded(gpu_model):
    """inference() consumes the CUDA generator like the reference: one randn(B,T,C) draw (diffusion.py:364)."""
    from said_b200.synth import synthetic_batch

    m = gpu_model("epsilon")
    wave = synthetic_batch(2, 1.0).to(DEV)
    torch.manual_seed(0)
    a = m.inference(wave, num_inference_steps=10, guidance_scale=2.0).result
    torch.manual_seed(0)
    noise = torch.randn(2, 60, 32, device=DEV)
    b = run(m, wave, noise, steps=10).result
    assert torch.equal(a, b)
    assert a.shape == (2, 60, 32) and a.device.type == "cuda"


def test_cpu_tensor_fails_loudly(gpu_model):
    from said_b200.synth import synthetic_batch

    with pytest.raises(RuntimeError, match="CUDA"):
        gpu_model("epsilon").inference(synthetic_batch(1, 1.0), num_inference_steps=2)


# ------------------------------------------------------------------------------------------------ DDPM (8(f) rank 2)
def test_ddpm_step_bit_exact(gpu_model):
    """Fused step kernel with scheduler code 1 == restated DDPMScheduler.step, bit for bit."""
    from said_b200 import scheduler as S

    eng = gpu_model()._engine(torch.device(DEV))
    g = torch.Generator().manual_seed(5)
    x = torch.randn(3, 60, 32, generator=g)
    e = torch.randn(3, 60, 32, generator=g)
    z = torch.randn(3, 60, 32, generator=g)
    for code, pt in enumerate(("epsilon", "sample", "v_prediction")):
        sch = S.DDPMScheduler(1000, beta_schedule="squaredcos_cap_v2", prediction_type=pt)
        sch.set_timesteps(50)
        for t in (980, 500, 0):
            row = S.ddpm_step_table(sch, [t])[0]
            out = eng.op_ddim_step(e.to(DEV), x.to(DEV), False, 1.0, 0.0, code, row, eta_noise=z.to(DEV), scheduler=1)
            want = sch.step(e, t, x, variance_noise=z).prev_sample
            assert np.array_equal(out.cpu().numpy(), want.numpy()), (pt, t, maxdiff(out, want))


def test_ddpm_chain_vs_oracle(state_dict):
    """SAID_UNet1D(noise_scheduler=DDPMScheduler): 1 s clip, 10 ancestral steps, CFG 2.0, against the CPU oracle fed the same
    initial and per-step noise (the draws come from the device generator in the reference's order).  Tolerance 5e-4."""
    from oracle import said_oracle as O
    from said_b200 import scheduler as S
    from said_b200.model.diffusion import SAID_UNet1D
    from said_b200.synth import synthetic_batch

    m = SAID_UNet1D(noise_scheduler=S.DDPMScheduler, prediction_type="epsilon")
    m.load_state_dict(state_dict)
    m.to(DEV).eval()
    wave = synthetic_batch(2, 1.0)
    torch.manual_seed(11)
    with torch.no_grad():
        out = m.inference(wave.to(DEV), num_inference_steps=10, guidance_scale=2.0)
    torch.manual_seed(11)
    noise = torch.randn(2, 60, 32, device=DEV).cpu()
    steps = [torch.randn(2, 60, 32, device=DEV).cpu() for _ in range(9)] + [torch.zeros(2, 60, 32)]
    with torch.no_grad():
        ref, _ = O.inference(state_dict, wave, num_inference_steps=10, guidance_scale=2.0, noise=noise,
                             eta_noise=torch.stack(steps), scheduler="ddpm")
    e = maxdiff(out.result, ref)
    print("ddpm chain vs oracle", e)
    assert e < 5e-4


def test_repeated_clip_is_encoded_once(gpu_model):
    """script/test_inference.py:167-168 repeats one clip over the batch: the audio encoder then runs for one row
    (dedup_audio) and the result equals the one with de-duplication switched off up to encoder tile-shape rounding."""
    from said_b200.synth import synthetic_batch

    m = gpu_model("epsilon")
    eng = m._engine(torch.device(DEV))
    wave = synthetic_batch(1, 1.0).repeat(6, 1)
    g = torch.Generator().manual_seed(2)
    noise = torch.randn(6, 60, 32, generator=g)
    n0 = eng.launches
    a = run(m, wave, noise, steps=4).result.cpu()
    la = eng.launches - n0
    m.dedup_audio = False
    try:
        n0 = eng.launches
        b = run(m, wave, noise, steps=4).result.cpu()
        lb = eng.launches - n0
    finally:
        m.dedup_audio = True
    print("dedup", maxdiff(a, b), la, lb)
    assert maxdiff(a, b) < 1e-4
    assert not torch.equal(a[0], a[1])      # different noise per row: different results


def test_large_family_encoder_golden(golden_dir, large_family):
    """wav2vec2-large family on the CUDA path (per-conv bias + LayerNorm + GELU feature extractor, pre-LN transformer layers,
    final LayerNorm) against the output of the reference's own ModifiedWav2Vec2Model.  Reference fp32-vs-fp64 floor for this
    configuration: 1.7e-5; tolerance 1e-4."""
    from said_b200.model.diffusion import SAID_UNet1D

    cfg, sd = large_family
    gd = load(golden_dir, "audio_encoder_large_family_1s.npz")
    m = SAID_UNet1D(audio_config=cfg)
    m.load_state_dict(sd, strict=False)
    m.to(DEV).eval()
    with torch.no_grad():
        emb = m.get_audio_embedding(torch.from_numpy(gd["wave"]).to(DEV), 60)
    e32, e64 = maxdiff(emb, gd["emb"]), maxdiff(emb, gd["emb64"])
    print("large-family encoder", e32, e64)
    assert emb.shape == (1, 60, 256)
    assert e32 < 1e-4 and e64 < 1e-4
    # and a batch, through the full inference path (encoder -> K/V hoist -> 4 steps): finite, in range, per-clip independent
    wave = torch.from_numpy(gd["wave"]).repeat(3, 1)
    wave[1] = wave[1].flip(0)
    torch.manual_seed(0)
    with torch.no_grad():
        out = m.inference(wave.to(DEV), num_inference_steps=4, guidance_scale=2.0).result
    assert out.shape == (3, 60, 32) and bool(torch.isfinite(out).all())


def test_process_audio_device_matches_host(gpu_model):
    """Device-side per-utterance normalisation (8(f) rank 1) == the host feature extractor's, to fp32 rounding."""
    from said_b200.synth import synthetic_waveform

    m = gpu_model()
    raw = np.stack([synthetic_waveform(i, 1.0) * (1.0 + i) + 0.05 * i for i in range(3)]).astype(np.float32)
    want = m.process_audio(list(raw))
    got = m.process_audio_device(torch.from_numpy(raw).to(DEV))
    e = maxdiff(got, want)
    print("process_audio_device", e)
    assert got.device.type == "cuda" and got.shape == want.shape
    assert e < 2e-6


def test_gemm_tail_slivers_match_fp32(gpu_model):
    """150 row tiles on 148 SMs: the two leftover tiles of every N = 192 GEMM are cut into 16-column slivers spread over
    24 CTAs (gemm_tc.cuh "tail balancing").  One UNet forward of 320 samples x 60 frames (19 200 rows) in tf32x3 against
    the fp32 FFMA kernels of the same engine; tolerance 2e-4 (the precision-mode test's bound for block activations)."""
    m = gpu_model()
    eng = m._engine(torch.device(DEV))
    g = torch.Generator().manual_seed(21)
    x = torch.randn(320, 60, 32, generator=g).to(DEV)
    ctx = torch.randn(320, 60, 768, generator=g).to(DEV)
    t = torch.randint(0, 1000, (320,), generator=g)
    outs = {}
    for mode in ("tf32x3", "fp32"):
        eng.set_precision(mode, -1, "fp32")
        try:
            outs[mode] = eng.denoiser_forward(x, t, ctx).cpu()
        finally:
            eng.set_precision(m.precision, m.tc_min_rows, m.encoder_precision)
    e = maxdiff(outs["tf32x3"], outs["fp32"])
    print("tail slivers tf32x3 vs fp32", e)
    assert e < 2e-4


def test_full_size_forward_tc_vs_fp32(gpu_model):
    """BASELINE configs[2] shape: one UNet forward of 128 samples x 300 frames (38 400 rows = 300 / 900 / 2400 GEMM tiles on 148
    SMs: every tail-sliver width the production run uses, the 3-tile attention, the 4-CTA-cluster GroupNorm) in tf32x3 against
    the fp32 FFMA kernels of the same engine.  Tolerance 2e-4 (block-activation bound of the precision-mode test)."""
    m = gpu_model()
    eng = m._engine(torch.device(DEV))
    g = torch.Generator().manual_seed(33)
    x = torch.randn(128, 300, 32, generator=g).to(DEV)
    ctx = torch.randn(128, 300, 768, generator=g).to(DEV)
    t = torch.randint(0, 1000, (128,), generator=g)
    outs = {}
    for mode in ("tf32x3", "fp32"):
        eng.set_precision(mode, -1, "fp32")
        try:
            outs[mode] = eng.denoiser_forward(x, t, ctx).cpu()
        finally:
            eng.set_precision(m.precision, m.tc_min_rows, m.encoder_precision)
    e = maxdiff(outs["tf32x3"], outs["fp32"])
    print("full-size forward tf32x3 vs fp32", e)
    assert bool(torch.isfinite(outs["tf32x3"]).all())
    assert e < 2e-4
