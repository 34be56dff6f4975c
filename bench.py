#!/usr/bin/env python
"""Benchmark of the SAiD inference hot path on B200 (BASELINE.json metric: clips/sec and
denoise-steps/sec, 5 s @ 16 kHz clips, 1000-step DDIM loop ("1000-step DDPM" in BASELINE's wording:
what script/inference.py runs is DDIMScheduler with num_inference_steps=1000, eta=0), 1/2/4/8 B200).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]

One "step" = one pass of the hot path over one batch = one ``SAID_UNet1D.inference()`` call on
``batch`` clips per GPU (audio encoder + K/V hoist + 1000 x [UNet forward x 2 CFG branches + scheduler
step]).  Workload = BASELINE configs[2] (batch 64 x 5 s, 1000 steps, one B200), the per-GPU shard of
configs[3] (512 clips over 8 GPUs); weak scaling: every rank runs its own 64 clips, the only collective
is the gather of results.  The same JSON line also carries the other BASELINE configs as extra keys:
``latency_b1`` (configs[1]), ``config0_1s_10steps`` (configs[0]), ``config4_editing`` (configs[4], 16 clips per GPU),
``strong_b64`` (64 clips split over the N GPUs) and ``torch_eager_b200`` (the oracle's PyTorch-eager restatement run
on the same B200: the "reference on this GPU" comparator of SURVEY 8(d)).

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for the definition of every key.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

SECONDS = 5.0
NUM_STEPS = 1000
GUIDANCE = 2.0
FPS = 60
SR = 16000
CPU_BATCH = 8            # configs 3-5 are timed on the CPU at batch 8 and scaled linearly
ROUND = 2


# --------------------------------------------------------------------------------------------------
# algorithmic work; K/V hoisted, 3-key cross attention, 2 forwards per clip-step
# --------------------------------------------------------------------------------------------------
def denoiser_flops_per_sample_forward(T: int, c_in: int = 32):
    C, FF = 192, 768
    g = 0
    g += 2 * T * C * (3 * c_in)                                   # input conv
    g += 3 * (2 * T * C * 576 * 2)                                # 3 plain ResBlocks
    g += 2 * (2 * T * C * 1152 + 2 * T * C * 576 + 2 * T * C * 384)  # 2 concat ResBlocks (+1x1 skip)
    g += 4 * (2 * T * 3 * C * C + 2 * T * C * C)                  # self-attention projections
    g += 4 * (2 * T * C * C * 2)                                  # cross-attention q / out projections
    g += 4 * (2 * T * 2 * FF * C + 2 * T * C * FF)                # GEGLU FFN
    g += 4 * (2 * T * C * C)                                      # proj_out
    g += 2 * T * c_in * 576                                       # output conv
    attn = 4 * (2 * 2 * T * T * C)                                # QK^T and PV, 6 heads x 32
    xattn = 4 * (2 * 2 * 3 * T * C)
    temb = 2 * (192 * 768 + 768 * 768) + 5 * 2 * 768 * 192
    return {"gemm": g, "self_attention": attn, "cross_attention3": xattn, "time_embed": temb,
            "total": g + attn + xattn + temb}


def gemm_flops_executed_per_clip_step(T: int, c_in: int = 32) -> int:
    """GEMM FLOPs the engine actually executes per clip and loop iteration under CFG (2 branches).  Less than
    2 x denoiser_flops_per_sample_forward()["gemm"]: the input conv, the first ResBlock and the first block's q/k/v and
    attention out-projection run once for both branches (identical inputs), and the cross-attention q / out projections run
    for the conditional branch only (the null-condition branch's cross-attention is a constant).  The roofline numerator
    uses THIS figure, so work that is skipped is not credited."""
    C = 192
    full = 2 * denoiser_flops_per_sample_forward(T, c_in)["gemm"]
    shared = 2 * T * C * (3 * c_in) + 2 * T * C * 576 * 2 + 2 * T * 3 * C * C + 2 * T * C * C
    cond_only = 4 * (2 * T * C * C * 2)
    return full - shared - cond_only


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def cpu_model_name() -> str:
    try:
        with open("/proc/cpuinfo") as f:
            for ln in f:
                if ln.lower().startswith("model name"):
                    return ln.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
# CPU arm: the oracle (torch-CPU restatement of the reference's algorithm with the reference's cost
# structure) on the host cores.  The reference itself cannot travel to the GPU box (pure Python with
# missing third-party deps), so kind = "port".
# --------------------------------------------------------------------------------------------------
def cpu_sample(sd, batch: int, loop_iters: int, threads: int):
    """One bounded sample of the workload on the host: audio encoder for `batch` 5 s clips + `loop_iters` iterations of
    the 1000-step loop at that batch; returns (seconds_encoder, seconds_per_loop_iteration)."""
    from oracle import said_oracle as O
    from said_b200.synth import synthetic_batch

    torch.set_num_threads(threads)
    wave = synthetic_batch(batch, SECONDS)
    T = int(wave.shape[1] / SR * FPS)
    g = torch.Generator().manual_seed(0)
    noise = torch.randn(batch, T, 32, generator=g)
    with torch.no_grad():
        t0 = time.perf_counter()
        emb = O.audio_embedding(sd, wave, T)
        t1 = time.perf_counter()
        O.inference(sd, wave, num_inference_steps=NUM_STEPS, guidance_scale=GUIDANCE, noise=noise, audio_emb=emb,
                    step_limit=loop_iters)
        t2 = time.perf_counter()
    return t1 - t0, (t2 - t1) / loop_iters


def cpu_clips_per_s(batch: int, t_enc: float, t_iter: float) -> float:
    return batch / (t_enc + NUM_STEPS * t_iter)


def cpu_sample_text(batch, iters, threads, te, ti) -> str:
    return (f"oracle (torch-CPU restatement with the reference's cost structure: K/V re-projected every step, full T x T masked "
            f"cross-attention, per-row mask loop) on {threads} host threads of '{cpu_model_name()}': Wav2Vec2 encoder for "
            f"{batch} x 5 s clips ({te:.2f} s) + {iters} of the {NUM_STEPS} loop iterations at batch {batch} under CFG "
            f"({ti * 1000:.1f} ms each); clips/s = {batch} / (t_encoder + {NUM_STEPS} * t_iteration), i.e. scaled linearly in "
            f"the step count as BASELINE.md section 3 prescribes")


def run_reference_arm(args, out):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from said_b200.synth import synthetic_state_dict

    sd = synthetic_state_dict(0)
    threads = os.cpu_count() or 1
    iters = args.cpu_iters
    for _ in range(min(args.warmup, 2)):
        cpu_sample(sd, CPU_BATCH, max(2, iters // 8), threads)
    vals, t_all, last = [], 0.0, (0.0, 0.0)
    for _ in range(args.steps):
        t0 = time.perf_counter()
        te, ti = cpu_sample(sd, CPU_BATCH, iters, threads)
        t_all += time.perf_counter() - t0
        vals.append(cpu_clips_per_s(CPU_BATCH, te, ti))
        last = (te, ti)
    v = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": "clips/sec", "value": v, "unit": "clips/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * t_all / max(1, args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "denoise_clip_steps_per_s": v * NUM_STEPS,
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": v, "unit": "clips/s", "cores": threads, "kind": "port", "cpu": cpu_model_name(),
                         "sample": "per step: " + cpu_sample_text(CPU_BATCH, iters, threads, *last)},
        "e2e": {"value": v, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    out.write(json.dumps(line) + "\n")
    out.flush()


def workload_config(args, world):
    return {
        "workload": f"BASELINE configs[2]: batch={args.batch} x {SECONDS:g}s clips per GPU, {NUM_STEPS}-step DDIM (eta 0), "
                    f"CFG {GUIDANCE}, epsilon prediction, T=300 frames x 32 blendshapes; x{world} GPUs = configs[3] sharding",
        "batch_per_gpu": args.batch, "global_batch": args.batch * world, "seconds": SECONDS, "num_inference_steps": NUM_STEPS,
        "guidance_scale": GUIDANCE, "parallelism": f"clips sharded x{world}, gather of results only",
        "weights": "synthetic seeded (said_b200.synth), reference state-dict layout",
        "l2": "per-step activation working set (batch 64: ~450 MB) exceeds the 126 MB L2; no flush between iterations",
    }


# --------------------------------------------------------------------------------------------------
def protect_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write to file descriptor 1 behind Python's back (NCCL prints
    "NCCL version ..." there when NCCL_DEBUG is set in the environment), so fd 1 is pointed at stderr for the whole run
    and the JSON line goes to a private duplicate of the original stdout."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(real, "w")


def synthetic_coefficients(batch: int, frames: int, seed: int = 0) -> torch.Tensor:
    """Smooth blendshape-coefficient curves in [0, 0.75] (range of the reference's data/blendshape_coeffs.zip): init_samples
    for the editing workload (the reference's CSVs are not shipped to the GPU box)."""
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(frames, dtype=torch.float32)[None, :, None] / 60.0
    f = 0.3 + 2.5 * torch.rand(batch, 1, 32, generator=g)
    ph = 6.2832 * torch.rand(batch, 1, 32, generator=g)
    amp = 0.375 * torch.rand(batch, 1, 32, generator=g)
    return (amp * (1.0 + torch.sin(6.2832 * f * t + ph))).clamp(0.0, 0.75)


def main():
    out = protect_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="clips per GPU")
    ap.add_argument("--cpu-iters", type=int, default=20, help="loop iterations per CPU sample (at batch 8)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-b1", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the configs[0] / configs[4] / strong-scaling / eager legs")
    ap.add_argument("--precision", default=None, help="contraction precision of the denoiser GEMMs (default: the model's default; "
                                                      "see include/said_b200.h said_set_precision)")
    args = ap.parse_args()

    if args.impl == "reference":
        run_reference_arm(args, out)
        return

    import torch.distributed as dist

    from said_b200.model.diffusion import SAID_UNet1D
    from said_b200.parallel import gather_clips
    from said_b200.synth import normalise_waveform, synthetic_state_dict, synthetic_waveform

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)

    sd = synthetic_state_dict(0)
    model = SAID_UNet1D(prediction_type="epsilon")
    model.load_state_dict(sd)
    model.to(dev).eval()
    if args.precision:
        model.precision = args.precision
    precision = model.precision
    eng = model._engine(dev)

    B = args.batch
    gB = B * world
    T = int(SECONDS * FPS)
    lo = rank * B
    wave_host = torch.from_numpy(np.stack([normalise_waveform(synthetic_waveform(lo + i, SECONDS)) for i in range(B)])).pin_memory()
    wave_dev = wave_host.to(dev)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    noise = torch.randn(B, T, 32, device=dev, generator=gen)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def step_resident():
        with torch.no_grad():
            o = model._run(wave_dev, noise, None, None, NUM_STEPS, 1.0, GUIDANCE, 0.0, 0.0, T, False, False, None)
        return gather_clips(o.result, gB)

    def step_e2e():
        # the caller's view: pinned host waveforms in, the whole batch's coefficients back on the host of rank 0
        with torch.no_grad():
            w = wave_host.to(dev, non_blocking=True)
            o = model.inference(w, num_inference_steps=NUM_STEPS, guidance_scale=GUIDANCE)
            res = gather_clips(o.result, gB, dst=0)
            return res.cpu() if res is not None else None

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = eng.launches
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), eng.launches - l0

    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_total, launches = timed(step_resident, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms_total / args.steps
    value = gB / (ms_step / 1000.0)

    step_e2e()
    ms_e2e_total, _ = timed(step_e2e, args.steps)
    e2e_value = gB / (ms_e2e_total / args.steps / 1000.0)

    # ---- per-kernel-family device time for the roofline: a short un-graphed run with an event after every launch
    prof_steps = 20
    model.use_cuda_graph = False
    model._profile_loop = True
    with torch.no_grad():
        model._run(wave_dev, noise, None, None, NUM_STEPS, prof_steps / NUM_STEPS, GUIDANCE, 0.0, 0.0, T, False, False, None)
    prof = eng.profile_end()
    model._profile_loop = False
    model.use_cuda_graph = True
    fl = denoiser_flops_per_sample_forward(T)
    peaks = load_peaks()
    gemm_keys = [k for k in prof if k.startswith("gemm")]
    gemm_ms = sum(prof[k]["ms"] for k in gemm_keys)
    gemm_launches = sum(prof[k]["launches"] for k in gemm_keys)
    # (profiling starts after the audio encoder and the K/V hoist: loop kernels only)
    loop_ms = sum(v["ms"] for v in prof.values())
    gemm_flops_per_step = B * gemm_flops_executed_per_clip_step(T)          # executed, not algorithmic-as-written
    gemm_flops_algorithmic = 2 * B * fl["gemm"]
    achieved_tflops = (gemm_flops_per_step * prof_steps) / (gemm_ms / 1000.0) / 1e12 if gemm_ms > 0 else 0.0
    traffic, traffic_src = None, None
    for name in (f"r{ROUND}_ncu_traffic.json", "r1_ncu_traffic.json"):
        tpath = os.path.join(ROOT, "profiles", name)
        if os.path.exists(tpath):
            with open(tpath) as f:
                tj = json.load(f)
            if tj.get("precision", "tf32x3") == precision:
                traffic = tj["dram_bytes_per_launch"]
                traffic_src = f"profiles/{name} (DRAM bytes per launch, ncu dram__bytes capture of the same kernels, see the file)"
                break
    roofline = {
        "kernel": f"tcgen05 GEMM family ({precision}): every Linear/Conv1d of the UNet" if precision != "fp32"
                  else "gemm_simt_kernel (fp32 FFMA): every Linear/Conv1d of the UNet",
        "bound": "tensor", "achieved": achieved_tflops, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
        "frac": achieved_tflops / peaks["bf16_tflops_sustained"], "traffic": traffic, "traffic_source": traffic_src,
        "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({peaks['source']})",
        "flops_per_launch_avg": gemm_flops_per_step * prof_steps / max(1, gemm_launches),
        "avg_launch_ms": gemm_ms / max(1, gemm_launches),
        "gemm_launches_per_step": gemm_launches / prof_steps,
        "flops_executed_per_step": gemm_flops_per_step, "flops_algorithmic_per_step": gemm_flops_algorithmic,
        "share_of_step": gemm_ms / loop_ms if loop_ms > 0 else None,
        "family_ms_share": {k: (v["ms"] / loop_ms if loop_ms > 0 else None) for k, v in prof.items()},
        "family_ms_per_step": {k: v["ms"] / prof_steps for k, v in prof.items()},
        "note": f"precision mode {precision}; numerator = FLOPs EXECUTED (one multiply-add per weight per row actually computed: the "
                "CFG-shared prefix and the constant null-condition cross-attention are not credited; the extra passes of the "
                f"hi/lo operand split are not counted either); measured with one CUDA event per launch over {prof_steps} un-graphed "
                "loop iterations at the bench batch",
    }

    dtype_names = {"fp32": "f32", "tf32x3": "f32 (3xTF32 tensor-core split, fp32 accumulate)", "tf32": "tf32",
                   "fp16x3": "f32 (fp16 hi/lo tensor-core split, 3 passes, fp32 accumulate)"}
    line = {
        "metric": "clips/sec", "value": value, "unit": "clips/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": dtype_names.get(precision, precision),
        "data": "synthetic", "denoise_clip_steps_per_s": value * NUM_STEPS,
        "config": workload_config(args, world),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "clips/s", "h2d_bytes_per_step": int(wave_host.numel() * 4) * world,
                "d2h_bytes_per_step": int(gB * T * 32 * 4),
                "note": "pinned host waveforms -> device every step; results gathered to rank 0 and copied to its host"},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "algorithmic_tflops_whole_step": 2 * gB * fl["total"] * NUM_STEPS / (ms_step / 1000.0) / 1e12,
    }

    if not args.no_extras:
        # ---- BASELINE configs[4]: editing (init_samples + mask), 16 clips per GPU (128 over 8), 50 DDIM steps
        EB = 16
        init = synthetic_coefficients(EB, T, seed=100 + rank).to(dev)
        masks = {"inbetween": torch.zeros(EB, T, 32, device=dev), "blendshape": torch.zeros(EB, T, 32, device=dev)}
        masks["inbetween"][:, :100] = 1.0
        masks["inbetween"][:, 200:] = 1.0
        masks["blendshape"][:, :, :16] = 1.0
        ew, en = wave_dev[:EB].contiguous(), noise[:EB].contiguous()
        edit = {}
        for tag, strength in (("inbetween", 1.0), ("blendshape", 0.6)):
            def estep(tag=tag, strength=strength):
                with torch.no_grad():
                    o = model._run(ew, en, init, masks[tag], 50, strength, GUIDANCE, 0.0, 0.0, T, False, False, None)
                return gather_clips(o.result, EB * world)
            for _ in range(2):
                estep()
            ms_e, _ = timed(estep, 5)
            edit[f"{tag}_strength{strength}"] = {"clips_per_s": EB * world / (ms_e / 5 / 1000.0), "ms_per_call": ms_e / 5,
                                                 "loop_iterations": int(50 * strength)}
        line["config4_editing"] = {"workload": f"BASELINE configs[4]: editing, {EB} clips x 5 s per GPU ({EB * world} total), 50 DDIM "
                                               "steps, CFG 2.0, incl. audio encoder; weak-scaled like the headline", **edit}
        # ---- strong scaling: BASELINE configs[2]'s 64 clips split over the N GPUs
        if world > 1 and 64 % world == 0:
            sb = 64 // world
            sw, sn = wave_dev[:sb].contiguous(), noise[:sb].contiguous()

            def sstep():
                with torch.no_grad():
                    o = model._run(sw, sn, None, None, NUM_STEPS, 1.0, GUIDANCE, 0.0, 0.0, T, False, False, None)
                return gather_clips(o.result, 64)
            sstep()
            ms_s, _ = timed(sstep, 2)
            line["strong_b64"] = {"workload": f"64 clips x 5 s split over {world} GPUs ({sb} per GPU), 1000 steps",
                                  "clips_per_s": 64 / (ms_s / 2 / 1000.0), "ms_per_call": ms_s / 2}

    if rank == 0 and world == 1:
        if not args.no_b1:
            # BASELINE configs[1]: one clip, latency-bound
            w1, n1 = wave_dev[:1].contiguous(), noise[:1].contiguous()

            def b1():
                with torch.no_grad():
                    model._run(w1, n1, None, None, NUM_STEPS, 1.0, GUIDANCE, 0.0, 0.0, T, False, False, None)

            for _ in range(2):
                b1()
            ms1, _ = timed(b1, 3)
            line["latency_b1"] = {"workload": "BASELINE configs[1]: 1 x 5 s clip, 1000 steps", "clips_per_s": 1000.0 / (ms1 / 3),
                                  "ms_per_denoise_step": ms1 / 3 / NUM_STEPS, "ms_per_clip": ms1 / 3,
                                  "path": "engine default: 602 rows >= 512, tensor-core kernels with 32-column weight tile images"}
            # the same clip on the IEEE fp32 FFMA kernels (what runs below the row threshold)
            model.tc_min_rows = 1 << 30
            try:
                for _ in range(2):
                    b1()
                ms1t, _ = timed(b1, 3)
            finally:
                model.tc_min_rows = 0
                eng.set_precision(model.precision, -1, model.encoder_precision)
            line["latency_b1"]["ffma_kernels"] = {"ms_per_denoise_step": ms1t / 3 / NUM_STEPS, "clips_per_s": 1000.0 / (ms1t / 3)}
        if not args.no_extras:
            from oracle import said_oracle as O

            # ---- BASELINE configs[0]: 1 x 1 s, 10 DDIM steps, through the public API from a host waveform
            w0 = torch.from_numpy(normalise_waveform(synthetic_waveform(0, 1.0)))[None].pin_memory()

            def c0():
                with torch.no_grad():
                    torch.manual_seed(0)
                    return model.inference(w0.to(dev, non_blocking=True), num_inference_steps=10, guidance_scale=GUIDANCE).result.cpu()

            for _ in range(3):
                c0()
            ms0, _ = timed(c0, 20)
            torch.set_num_threads(os.cpu_count() or 1)
            g0 = torch.Generator().manual_seed(0)
            n0 = torch.randn(1, 60, 32, generator=g0)
            with torch.no_grad():
                O.inference(sd, w0, num_inference_steps=10, guidance_scale=GUIDANCE, noise=n0)
                t0 = time.perf_counter()
                for _ in range(3):
                    O.inference(sd, w0, num_inference_steps=10, guidance_scale=GUIDANCE, noise=n0)
                cpu0 = (time.perf_counter() - t0) / 3
            line["config0_1s_10steps"] = {
                "workload": "BASELINE configs[0]: 1 x 1 s clip, 10 DDIM steps, CFG 2.0, incl. audio encoder, host waveform in / host result out",
                "ms_per_call": ms0 / 20, "clips_per_s": 1000.0 / (ms0 / 20),
                "cpu_oracle_ms_per_call": cpu0 * 1000.0, "cpu_cores": os.cpu_count()}
            # ---- the caller pattern of script/test_inference.py:148-202 for one utterance: the clip repeated over the batch, 72
            #      generations in chunks of 64, every result cut to window_len frames and written to its own CSV
            import tempfile

            from said_b200.util.audio import fit_audio_unet
            from said_b200.util.blendshape import DEFAULT_BLENDSHAPE_CLASSES, save_blendshape_coeffs, save_blendshape_coeffs_batch

            raw = torch.from_numpy(synthetic_waveform(0, SECONDS))[: int(SECONDS * SR) - 123]      # not a multiple of the 800-sample hop
            with torch.no_grad(), tempfile.TemporaryDirectory() as td:
                def utterance(writer):
                    t0 = time.perf_counter()
                    fit = fit_audio_unet(raw.to(dev), SR, FPS, 1)
                    wp = model.process_audio_device(fit.waveform)
                    wb = wp.repeat(64, 1)
                    torch.manual_seed(0)
                    chunks = []
                    for cs in (64, 8):
                        o = model.inference(waveform_processed=wb[:cs], num_inference_steps=NUM_STEPS, guidance_scale=GUIDANCE)
                        chunks.append(o.result[:, : fit.window_size].cpu().numpy())
                    t1 = time.perf_counter()
                    res = np.concatenate(chunks, 0)
                    paths = [os.path.join(td, f"u-{i}.csv") for i in range(res.shape[0])]
                    if writer == "batched":
                        save_blendshape_coeffs_batch(res, DEFAULT_BLENDSHAPE_CLASSES, paths)
                    else:
                        for i, pth in enumerate(paths):
                            save_blendshape_coeffs(res[i], DEFAULT_BLENDSHAPE_CLASSES, pth)
                    return t1 - t0, time.perf_counter() - t1
                utterance("batched")
                g_b, h_b = utterance("batched")
                g_p, h_p = utterance("pandas")
            line["caller_test_inference"] = {
                "workload": "script/test_inference.py:148-202 for one 5 s utterance: clip repeated x64 (encoded once), 72 generations in chunks "
                            "of 64 + 8, 1000 steps, results cut to window_len and written as 72 CSV files",
                "gpu_s": g_b, "csv_s_batched_writer": h_b, "csv_s_pandas_per_file": h_p, "generations_per_s": 72 / (g_b + h_b)}
            # ---- the reference's algorithm in PyTorch eager ON THIS B200 (oracle with device tensors): SURVEY 8(d) comparator
            try:
                sdg = {k: v.to(dev) for k, v in sd.items()}
                eager = {}
                for eb, iters in ((1, 10), (64, 3)):
                    w = wave_dev[:eb].contiguous()
                    nz = noise[:eb].contiguous()
                    with torch.no_grad():
                        torch.cuda.synchronize(dev)
                        t0 = time.perf_counter()
                        emb = O.audio_embedding(sdg, w, T)
                        torch.cuda.synchronize(dev)
                        t_enc = time.perf_counter() - t0
                        O.inference(sdg, w, num_inference_steps=NUM_STEPS, guidance_scale=GUIDANCE, noise=nz, audio_emb=emb, step_limit=2)
                        torch.cuda.synchronize(dev)
                        t0 = time.perf_counter()
                        O.inference(sdg, w, num_inference_steps=NUM_STEPS, guidance_scale=GUIDANCE, noise=nz, audio_emb=emb, step_limit=iters)
                        torch.cuda.synchronize(dev)
                        t_it = (time.perf_counter() - t0) / iters
                    eager[f"batch{eb}"] = {"ms_per_denoise_step": t_it * 1000.0, "encoder_s": t_enc,
                                           "clips_per_s_extrapolated": eb / (t_enc + NUM_STEPS * t_it)}
                    del emb
                del sdg
                torch.cuda.empty_cache()
                line["torch_eager_b200"] = {
                    "what": "oracle/said_oracle.py (the reference's algorithm and op structure in plain PyTorch: cuDNN/cuBLAS/ATen "
                            "kernels, per-row mask loop, K/V re-projected every step) with all tensors on this GPU; wall clock with "
                            "synchronisation, a few loop iterations extrapolated to 1000", **eager}
            except Exception as ex:  # the comparator must never take the bench line down
                line["torch_eager_b200"] = {"error": repr(ex)[:300]}
        if not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            te, ti = cpu_sample(sd, CPU_BATCH, args.cpu_iters, threads)
            line["cpu_baseline"] = {
                "value": cpu_clips_per_s(CPU_BATCH, te, ti), "unit": "clips/s", "cores": threads, "kind": "port",
                "cpu": cpu_model_name(), "sample": cpu_sample_text(CPU_BATCH, args.cpu_iters, threads, te, ti),
            }
    if rank == 0:
        out.write(json.dumps(line) + "\n")
        out.flush()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
