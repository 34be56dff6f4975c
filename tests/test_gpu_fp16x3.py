"""GPU tests of the fp16x3 path: the TMA-fed tcgen05 kind::f16 GEMM over fp16 hi/lo operand pairs (gemm_h.cuh), the kernels that
produce pair-format operands, and the denoiser forward / step loop built from them -- against fp64 torch math, the golden
vectors of the unmodified reference and the CPU oracle."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def maxdiff(a, b):
    return float((torch.as_tensor(a).detach().cpu().double() - torch.as_tensor(b).detach().cpu().double()).abs().max())


def _engine(gpu_model):
    return gpu_model()._engine(torch.device(DEV))


@pytest.mark.parametrize("M,Cin,N", [(128, 192, 192), (300, 192, 192), (1000, 64, 192), (4096, 768, 192), (2000, 192, 576),
                                     (1500, 192, 1536), (19264, 192, 192), (38528, 192, 192), (700, 192, 32)])
def test_gemm_h_plain_vs_fp64(gpu_model, M, Cin, N):
    """out = A W + bias on the tensor cores with fp16 hi/lo operands (3 passes) against fp64: the error of an fp32 GEMM, not of
    an fp16 one.  19264 / 38528 rows = 151 / 301 row tiles on 148 SMs: the leftover tiles run as N-slivers."""
    eng = _engine(gpu_model)
    g = torch.Generator().manual_seed(M + Cin + N)
    a = torch.randn(M, Cin, generator=g) * (1.0 + 3.0 * torch.rand(M, 1, generator=g))
    w = torch.randn(Cin, N, generator=g) / np.sqrt(Cin)
    bias = torch.randn(N, generator=g)
    out = eng.op_gemm_h(a.to(DEV), w, 1, bias.to(DEV)).cpu()
    ref = a.double() @ w.double() + bias.double()
    err = maxdiff(out, ref)
    scale = float(ref.abs().max())
    print(f"gemm_h plain M={M} K={Cin} N={N}: max err {err:.3e} (|ref| max {scale:.2f})")
    assert err < 8e-6 * max(1.0, scale)


@pytest.mark.parametrize("M,Cin,N", [(602, 192, 192), (1204, 384, 192), (38528, 192, 192), (903, 192, 32)])
def test_gemm_h_conv3_vs_fp64(gpu_model, M, Cin, N):
    """Conv1d(k=3, pad=1) as three row-shifted TMA boxes: rows m-1, m, m+1 (rows outside the tensor read as zero)."""
    eng = _engine(gpu_model)
    g = torch.Generator().manual_seed(7 * M + Cin + N)
    a = torch.randn(M, Cin, generator=g)
    w = torch.randn(3 * Cin, N, generator=g) / np.sqrt(3 * Cin)
    out = eng.op_gemm_h(a.to(DEV), w, 3, None).cpu()
    ad = a.double()
    z = torch.zeros(1, Cin, dtype=torch.float64)
    prev, nxt = torch.cat([z, ad[:-1]]), torch.cat([ad[1:], z])
    wd = w.double()
    ref = prev @ wd[:Cin] + ad @ wd[Cin:2 * Cin] + nxt @ wd[2 * Cin:]
    err = maxdiff(out, ref)
    print(f"gemm_h conv3 M={M} Cin={Cin} N={N}: max err {err:.3e}")
    assert err < 8e-6 * max(1.0, float(ref.abs().max()))


def test_gemm_h_small_and_large_magnitudes(gpu_model):
    """The pair format keeps ~22 bits from 2^-3 up to fp16's range and an absolute error of 2^-25 below; weights are pre-scaled
    by a power of two per matrix, so tiny weights lose nothing."""
    eng = _engine(gpu_model)
    g = torch.Generator().manual_seed(3)
    a = torch.randn(512, 192, generator=g)
    a[:128] *= 1e-3
    a[128:256] *= 300.0
    w = torch.randn(192, 192, generator=g) * 1e-4
    out = eng.op_gemm_h(a.to(DEV), w, 1, None).cpu()
    ref = a.double() @ w.double()
    rel = ((out.double() - ref).abs().amax(dim=1) / ref.abs().amax(dim=1))
    print("gemm_h magnitudes: rel err small rows", float(rel[:128].max()), "large rows", float(rel[128:256].max()), "unit rows", float(rel[256:].max()))
    assert float(rel[128:].max()) < 4e-6
    assert float(rel[:128].max()) < 1e-4      # |a| ~ 1e-3: the lo plane is an fp16 subnormal, absolute error 2^-25


def test_fp16_overflow_is_reported(gpu_model):
    eng = _engine(gpu_model)
    a = torch.ones(128, 64)
    a[5, 7] = 1.0e5
    from said_b200._lib import SaidLibraryError

    eng.op_gemm_h(a.to(DEV), torch.ones(64, 192), 1, None)
    with pytest.raises(SaidLibraryError):
        eng.check_status()
    eng.check_status()      # cleared


@pytest.mark.parametrize("T,B", [(300, 2), (60, 3), (128, 1), (129, 1), (257, 2), (304, 1), (17, 2), (469, 2), (512, 1)])
def test_self_attention_h_vs_fp64(gpu_model, T, B):
    """Flash-style tcgen05 attention over fp16 hi/lo pairs (online softmax, probabilities kept in tensor memory) against fp64
    math, including the longest sequence in the reference's data (469 frames); tolerance 2e-5 like the 3xTF32 kernel."""
    eng = _engine(gpu_model)
    heads, hd = 6, 32
    g = torch.Generator().manual_seed(T)
    qkv = torch.randn(B, T, 3 * heads * hd, generator=g)
    qkv[:, :, : heads * hd] *= 1.5        # sharper softmax than unit-variance q k: exercises the running-max rescale
    qkv = qkv.to(DEV)
    out = eng.op_self_attention_h(qkv, heads)
    q, k, v = qkv.double().chunk(3, dim=-1)
    sh = lambda t: t.reshape(B, T, heads, hd).transpose(1, 2)  # noqa: E731
    ref = torch.softmax(sh(q) @ sh(k).transpose(-1, -2) * hd**-0.5, dim=-1) @ sh(v)
    ref = ref.transpose(1, 2).reshape(B, T, heads * hd)
    err = maxdiff(out, ref)
    print("fp16x3 attention", T, err)
    assert err < 2e-5


TAP_ORDER = ["input_blocks.0", "input_blocks.1.0", "input_blocks.1.1", "middle_block.0", "middle_block.1",
             "middle_block.2", "output_blocks.0.0", "output_blocks.0.1", "output_blocks.1.0", "output_blocks.1.1"]


def test_denoiser_forward_fp16x3_golden(golden_dir, gpu_model):
    """One UNet forward on the fp16x3 path (forced on for this small problem) against the reference's own block activations
    and output.  Same tolerance as the 3xTF32 mode (2e-4 on activations; the fp32 kernels: 2e-5)."""
    gd = np.load(os.path.join(golden_dir, "denoiser_forward.npz"))
    m = gpu_model()
    eng = m._engine(torch.device(DEV))
    eng.set_precision("fp16x3", 1, "fp32")
    try:
        out, taps = eng.denoiser_forward(torch.from_numpy(gd["x"]).to(DEV), torch.from_numpy(gd["t"]),
                                         torch.from_numpy(gd["ctx"]).to(DEV), taps=True)
    finally:
        eng.set_precision(m.precision, -1, m.encoder_precision)
    errs = {name: maxdiff(taps[i].transpose(1, 2), gd["act_" + name]) for i, name in enumerate(TAP_ORDER) if "act_" + name in gd.files}
    errs["out"] = maxdiff(out, gd["y64"])
    print("fp16x3", errs)
    assert all(v < 2e-4 for v in errs.values()), errs


def _run(m, wave, noise, init=None, mask=None, steps=10, strength=1.0, gs=2.0):
    wave = torch.as_tensor(wave).to(DEV)
    T = int(wave.shape[1] / 16000 * 60)
    cv = lambda a: None if a is None else torch.as_tensor(a).to(DEV)  # noqa: E731
    with torch.no_grad():
        return m._run(wave, cv(noise), cv(init), cv(mask), steps, strength, gs, 0.0, 0.0, T, False, False, None, return_latents=True)


class _Mode:
    def __init__(self, m, precision, tc_min_rows=0):
        self.m, self.p, self.r = m, precision, tc_min_rows

    def __enter__(self):
        self.saved = (self.m.precision, self.m.tc_min_rows)
        self.m.precision, self.m.tc_min_rows = self.p, self.r
        return self.m

    def __exit__(self, *a):
        self.m.precision, self.m.tc_min_rows = self.saved
        self.m._engine(torch.device(DEV)).set_precision(self.m.precision, -1, self.m.encoder_precision)


def test_chain_1000_steps_fp16x3(golden_dir, gpu_model):
    """The 1000-step free-running v-prediction chain with every denoiser contraction on the fp16x3 path (forced on for this single
    clip) against the reference: the stated tolerance of the tensor-core modes, 1e-3 (3xTF32 measures 2.6e-4)."""
    from said_b200.synth import normalise_waveform, synthetic_waveform

    gd = np.load(os.path.join(golden_dir, "chain_5s_1000steps_v_prediction.npz"))
    wave = torch.from_numpy(normalise_waveform(synthetic_waveform(0, 5.0)))[None]
    with _Mode(gpu_model("v_prediction"), "fp16x3", 1) as m:
        out = _run(m, wave, gd["noise"], steps=1000)
    e32, e64 = maxdiff(out.result, gd["result"]), maxdiff(out.result, gd["result64"])
    print("fp16x3 1000-step chain", e32, e64)
    assert e32 < 1e-3 and e64 < 1e-3


def test_bench_shape_chain_fp16x3_vs_oracle(gpu_model, state_dict):
    """32 clips x 5 s, CFG with the shared guidance prefix, 4 DDIM steps on the fp16x3 path against the CPU oracle (two clips)
    and against the 3xTF32 mode of the same engine."""
    from oracle import said_oracle as O
    from said_b200.synth import synthetic_batch

    B = 32
    wave = synthetic_batch(B, 5.0)
    g = torch.Generator().manual_seed(17)
    noise = torch.randn(B, 300, 32, generator=g)
    with _Mode(gpu_model("epsilon"), "fp16x3") as m:
        out = _run(m, wave, noise, steps=4)
    res, lat = out.result.cpu(), out.latents.cpu()
    with _Mode(gpu_model("epsilon"), "tf32x3") as m:
        lat_tf = _run(m, wave, noise, steps=4).latents.cpu()
    sel = [0, B - 1]
    with torch.no_grad():
        ref, pre = O.inference(state_dict, wave[sel], num_inference_steps=4, guidance_scale=2.0, noise=noise[sel], return_preclamp=True)
    e, ep, em = maxdiff(res[sel], ref), maxdiff(lat[sel], pre[-1]), maxdiff(lat, lat_tf)
    print("fp16x3 bench-shape chain vs oracle: result", e, "pre-clamp", ep, "vs tf32x3 (all clips)", em)
    assert bool(torch.isfinite(res).all())
    assert e < 5e-4 and ep < 1e-3 and em < 1e-3


def test_editing_fp16x3_and_mismatched_init_length(gpu_model, state_dict):
    """Editing on the fp16x3 path at batch scale, with init_samples two frames SHORTER than the audio window (a previous result
    cut to floor(len * 60 / 16000) frames used as --init_sample_path: the reference keeps the init length for the latents and
    aligns the cross-attention with its general window, attention.py:170-189).  Also checked on the fp32 kernels (one clip)."""
    from oracle import said_oracle as O
    from said_b200.synth import synthetic_batch

    m = gpu_model("epsilon")
    B, T = 12, 298
    wave = synthetic_batch(B, 5.0)                       # audio window: 300 frames
    g = torch.Generator().manual_seed(31)
    t = torch.arange(T, dtype=torch.float32)[None, :, None] / 60.0
    init = (0.3 * (1.0 + torch.sin(6.2832 * (0.5 + torch.rand(B, 1, 32, generator=g)) * t))).clamp(0, 0.75)
    noise = torch.randn(B, T, 32, generator=g)
    mask = torch.zeros(B, T, 32)
    mask[:, :90] = 1.0
    mask[:, 210:] = 1.0
    sel = [0, B - 1]
    with torch.no_grad():
        ref, _ = O.inference(state_dict, wave[sel], init_samples=init[sel], mask=mask[sel], num_inference_steps=20,
                             guidance_scale=2.0, noise=noise[sel])
    for mode in ("fp16x3", "fp32"):
        with _Mode(m, mode, 1):
            res = _run(m, wave, noise, init=init, mask=mask, steps=20).result.cpu()
        assert res.shape == (B, T, 32)
        kept = mask.bool()
        assert torch.equal(res[kept], init.clamp(0, 1)[kept]), mode
        e = maxdiff(res[sel], ref)
        print("editing, init 298 frames vs 300-frame audio window:", mode, e)
        assert e < 2e-3


def test_audio_encoder_fp16x3(golden_dir, gpu_model):
    """Wav2Vec2 encoder with every dense contraction on the fp16x3 tensor-core path (forced on for this 1 s clip) against the
    reference's (transformers) output, and a batch of 8 clips against the fp32 kernels of the same engine."""
    from said_b200.synth import synthetic_batch

    gd = np.load(os.path.join(golden_dir, "audio_encoder_1s.npz"))
    m = gpu_model()
    saved = (m.encoder_precision, m.tc_min_rows)
    try:
        m.encoder_precision, m.tc_min_rows = "fp16x3", 1
        emb = m.get_audio_embedding(torch.from_numpy(gd["wave"]).to(DEV), 60)
        wave = synthetic_batch(8, 1.0).to(DEV)
        got = m.get_audio_embedding(wave, 60)
        m.encoder_precision = "fp32"
        ref = m.get_audio_embedding(wave, 60)
    finally:
        m.encoder_precision, m.tc_min_rows = saved
        m._engine(torch.device(DEV)).set_precision(m.precision, -1, m.encoder_precision)
    e32, e64, eb = maxdiff(emb, gd["emb"]), maxdiff(emb, gd["emb64"]), maxdiff(got, ref)
    print("fp16x3 encoder vs reference", e32, e64, "batch of 8 vs fp32 kernels", eb)
    assert e32 < 2e-4 and e64 < 2e-4 and eb < 2e-4


def test_large_family_encoder_fp16x3(golden_dir, large_family):
    """wav2vec2-large family (LayerNorm feature extractor with conv bias, pre-LN layers) on the fp16x3 path."""
    from said_b200.model.diffusion import SAID_UNet1D

    cfg, sd = large_family
    gd = np.load(os.path.join(golden_dir, "audio_encoder_large_family_1s.npz"))
    m = SAID_UNet1D(audio_config=cfg)
    m.load_state_dict(sd, strict=False)
    m.to(DEV).eval()
    m.encoder_precision, m.tc_min_rows = "fp16x3", 1
    with torch.no_grad():
        emb = m.get_audio_embedding(torch.from_numpy(gd["wave"]).to(DEV), 60)
    e32, e64 = maxdiff(emb, gd["emb"]), maxdiff(emb, gd["emb64"])
    print("fp16x3 large-family encoder", e32, e64)
    assert e32 < 2e-4 and e64 < 2e-4


@pytest.mark.gpu
@pytest.mark.parametrize("M", [1000, 19264, 38528])
def test_fused_feed_forward_vs_fp64(gpu_model, M):
    """The fused GEGLU feed-forward kernel against fp64 math (reference ``ldm/attention.py:25-51, 232-234`` as folded at load):
    M = 1000 (8 whole tiles, the last one partial), 19264 (151 tiles: one round + 3 leftover tiles split across CTAs) and 38528
    (the bench's 301 tiles: two rounds + 5 split tiles, reduced in a fixed order)."""
    eng = _engine(gpu_model)
    g = torch.Generator(device="cpu").manual_seed(7 + M)
    ln = torch.randn(M, 192, generator=g)
    x2 = torch.randn(M, 192, generator=g) * 2.0
    res = torch.randn(M, 192, generator=g)
    w1 = torch.randn(192, 1536, generator=g) / 192 ** 0.5
    b1 = torch.randn(1536, generator=g) * 0.1
    w2 = torch.randn(960, 192, generator=g) / 960 ** 0.5
    b2 = torch.randn(192, generator=g) * 0.1
    dev = eng.device
    out = eng.op_ffn_h(ln.to(dev), x2.to(dev), res.to(dev), w1, b1.to(dev), w2, b2.to(dev)).cpu().double()
    # fp64 reference on a sample of rows (all rows of the split tiles' region included)
    rows = torch.unique(torch.cat([torch.arange(0, min(M, 300)), torch.arange(max(0, M - 700), M), torch.randint(0, M, (400,), generator=g)]))
    a = ln[rows].double() @ w1.double() + b1.double()
    val, gate = a[:, 0::2], a[:, 1::2]
    ff = val * (0.5 * gate * (1.0 + torch.erf(gate / 2 ** 0.5)))
    ref = torch.cat([ff, x2[rows].double()], dim=1) @ w2.double() + b2.double() + res[rows].double()
    err = float((out[rows] - ref).abs().max())
    print(f"fused feed-forward M={M}: max err {err:.3e} (|ref| max {float(ref.abs().max()):.2f})")
    assert err < 4e-5
    # run to run bit-identical (the split tiles are reduced in a fixed order)
    out2 = eng.op_ffn_h(ln.to(dev), x2.to(dev), res.to(dev), w1, b1.to(dev), w2, b2.to(dev)).cpu().double()
    assert torch.equal(out, out2)


@pytest.mark.gpu
@pytest.mark.parametrize("B", [1, 2, 6, 12])
def test_small_batch_default_path_vs_oracle(gpu_model, state_dict, B):
    """One, two and six 5 s clips at DEFAULT settings: 602 / 1204 / 3612 denoiser rows are above the fp16x3 row threshold (512),
    so the tensor-core kernels run with the 32-column (1, 2 clips) or 64-column (6 clips) weight tile images, the two-GEMM
    feed-forward, one head per attention CTA and programmatic dependent launch.
    4 DDIM steps under CFG against the CPU oracle; and the same call on the FFMA kernels must differ in the last bits (proof that
    the default did not take them)."""
    from oracle import said_oracle as O
    from said_b200.synth import synthetic_batch

    wave = synthetic_batch(B, 5.0)
    g = torch.Generator().manual_seed(23 + B)
    noise = torch.randn(B, 300, 32, generator=g)
    m = gpu_model("epsilon")
    eng = m._engine(torch.device(DEV))
    eng.set_precision(m.precision, -1, m.encoder_precision)
    out = _run(m, wave, noise, steps=4)
    res, lat = out.result.cpu(), out.latents.cpu()
    m.tc_min_rows = 1 << 30
    try:
        lat_ffma = _run(m, wave, noise, steps=4).latents.cpu()
    finally:
        m.tc_min_rows = 0
        eng.set_precision(m.precision, -1, m.encoder_precision)
    with torch.no_grad():
        ref, pre = O.inference(state_dict, wave, num_inference_steps=4, guidance_scale=2.0, noise=noise, return_preclamp=True)
    e, ep, ef = maxdiff(res, ref), maxdiff(lat, pre[-1]), maxdiff(lat, lat_ffma)
    print(f"small batch {B}: result vs oracle {e:.3e}, pre-clamp {ep:.3e}, vs FFMA kernels {ef:.3e}")
    assert e < 5e-4 and ep < 1e-3
    assert 0.0 < ef < 1e-3


@pytest.mark.gpu
@pytest.mark.parametrize("B,seconds", [(3, 5.0), (5, 2.5), (7, 5.0), (9, 5.0), (13, 5.0), (21, 5.0), (27, 5.0), (33, 5.0), (41, 4.0), (2, 7.8), (2, 9.0), (20, 9.0)])
def test_shape_sweep_default_vs_fp32_kernels(gpu_model, B, seconds):
    """Odd batch sizes and clip lengths through every dispatch regime of the default path (32- / 64- / 96- / 192-column weight
    tiles, two-GEMM and fused feed-forward with 0..n split leftover tiles, one head / one query tile per attention CTA, PDL, the
    two-half-batch schedule with unequal halves) against the IEEE fp32 FFMA kernels of the same engine: 3 DDIM steps under CFG."""
    from said_b200.synth import synthetic_batch

    wave = synthetic_batch(B, seconds)
    T = int(wave.shape[1] / 16000 * 60)
    g = torch.Generator().manual_seed(100 + B)
    noise = torch.randn(B, T, 32, generator=g)
    m = gpu_model("epsilon")
    eng = m._engine(torch.device(DEV))
    eng.set_precision(m.precision, -1, m.encoder_precision)
    out = _run(m, wave, noise, steps=3)
    lat = out.latents.cpu()
    m.tc_min_rows = 1 << 30
    try:
        lat_ffma = _run(m, wave, noise, steps=3).latents.cpu()
    finally:
        m.tc_min_rows = 0
        eng.set_precision(m.precision, -1, m.encoder_precision)
    e = maxdiff(lat, lat_ffma)
    print(f"shape sweep B={B} T={T}: default vs FFMA kernels {e:.3e}")
    assert bool(torch.isfinite(lat).all())
    assert e < 1e-3
