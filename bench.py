#!/usr/bin/env python
"""Benchmark of the SAiD inference hot path on B200 (BASELINE.json metric: clips/sec and
denoise-steps/sec, 5 s @ 16 kHz clips, 1000-step DDIM loop ("1000-step DDPM" in BASELINE's wording:
what script/inference.py runs is DDIMScheduler with num_inference_steps=1000, eta=0), 1/2/4/8 B200).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]

One "step" = one pass of the hot path over one batch = one ``SAID_UNet1D.inference()`` call on
``batch`` clips per GPU (audio encoder + K/V hoist + 1000 x [UNet forward x 2 CFG branches + scheduler
step]).  Workload = BASELINE configs[2] (batch 64 x 5 s, 1000 steps, one B200), the per-GPU shard of
configs[3] (512 clips over 8 GPUs); weak scaling: every rank runs its own 64 clips, the only collective
is the all-gather of results.  The line also carries configs[1] (batch 1) as ``latency_b1``.

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


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


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
def cpu_sample(sd, loop_iters: int, threads: int):
    """One bounded sample of the workload on the host: audio encoder for one 5 s clip + `loop_iters`
    iterations of the 1000-step loop; returns (seconds_encoder, seconds_per_loop_iteration)."""
    from oracle import said_oracle as O
    from said_b200.synth import synthetic_batch

    torch.set_num_threads(threads)
    wave = synthetic_batch(1, SECONDS)
    T = int(wave.shape[1] / SR * FPS)
    g = torch.Generator().manual_seed(0)
    noise = torch.randn(1, T, 32, generator=g)
    with torch.no_grad():
        t0 = time.perf_counter()
        emb = O.audio_embedding(sd, wave, T)
        t1 = time.perf_counter()
        O.inference(sd, wave, num_inference_steps=NUM_STEPS, guidance_scale=GUIDANCE, noise=noise, audio_emb=emb,
                    step_limit=loop_iters)
        t2 = time.perf_counter()
    return t1 - t0, (t2 - t1) / loop_iters


def cpu_clips_per_s(t_enc: float, t_iter: float) -> float:
    return 1.0 / (t_enc + NUM_STEPS * t_iter)


def run_reference_arm(args, out):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from said_b200.synth import synthetic_state_dict

    sd = synthetic_state_dict(0)
    threads = os.cpu_count() or 1
    iters = args.cpu_iters
    for _ in range(args.warmup):
        cpu_sample(sd, max(2, iters // 10), threads)
    vals, t_all = [], 0.0
    for _ in range(args.steps):
        t0 = time.perf_counter()
        te, ti = cpu_sample(sd, iters, threads)
        t_all += time.perf_counter() - t0
        vals.append(cpu_clips_per_s(te, ti))
    v = float(np.mean(vals))
    sample = (f"per step: Wav2Vec2 encoder for one 5 s clip + {iters} of the {NUM_STEPS} loop iterations at batch 1 (CFG: 2 "
              f"UNet forwards per iteration); clips/s = 1 / (t_encoder + {NUM_STEPS} * t_iteration); per-clip cost is flat in "
              "batch on CPU (BASELINE.md section 2), so the same figure stands for the batch-64 workload")
    line = {
        "impl": "reference", "metric": "clips/sec", "value": v, "unit": "clips/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * t_all / max(1, args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "denoise_clip_steps_per_s": v * NUM_STEPS,
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": v, "unit": "clips/s", "cores": threads, "kind": "port", "sample": sample},
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
        "guidance_scale": GUIDANCE, "parallelism": f"clips sharded x{world}, all-gather of results only",
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


def main():
    out = protect_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="clips per GPU")
    ap.add_argument("--cpu-iters", type=int, default=100, help="loop iterations per CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-b1", action="store_true")
    ap.add_argument("--precision", default="tf32x3", choices=["fp32", "tf32x3", "tf32"],
                    help="contraction precision of the denoiser GEMMs (see include/said_b200.h said_set_precision)")
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
    model.precision = args.precision
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
            out = model._run(wave_dev, noise, None, None, NUM_STEPS, 1.0, GUIDANCE, 0.0, 0.0, T, False, False, None)
        return gather_clips(out.result, gB)

    def step_e2e():
        with torch.no_grad():
            w = wave_host.to(dev, non_blocking=True)
            out = model.inference(w, num_inference_steps=NUM_STEPS, guidance_scale=GUIDANCE)
            res = gather_clips(out.result, gB)
            return res.cpu()

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
    gemm_ms = sum(prof[k]["ms"] for k in ("gemm_conv3", "gemm_layernorm", "gemm_plain"))
    gemm_launches = sum(prof[k]["launches"] for k in ("gemm_conv3", "gemm_layernorm", "gemm_plain"))
    # (profiling starts after the audio encoder and the K/V hoist: loop kernels only)
    loop_ms = sum(v["ms"] for v in prof.values())
    gemm_flops_per_step = 2 * B * fl["gemm"]
    achieved_tflops = (gemm_flops_per_step * prof_steps) / (gemm_ms / 1000.0) / 1e12 if gemm_ms > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r1_ncu_traffic.json")
    if args.precision == "tf32x3" and os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f)["dram_bytes_per_launch"]
    roofline = {
        "kernel": ("gemm_tc_kernel (tcgen05, " + args.precision + ")" if args.precision != "fp32" else "gemm_simt_kernel (fp32 FFMA)")
                  + ": all loader/epilogue instantiations = every Linear/Conv1d of the UNet",
        "bound": "tensor", "achieved": achieved_tflops, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
        "frac": achieved_tflops / peaks["bf16_tflops_sustained"], "traffic": traffic,
        "traffic_source": "profiles/r1_ncu_traffic.json (bytes per launch, ncu --set full capture of the same kernels)" if traffic else None,
        "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({peaks['source']})",
        "flops_per_launch_avg": gemm_flops_per_step * prof_steps / max(1, gemm_launches),
        "avg_launch_ms": gemm_ms / max(1, gemm_launches),
        "share_of_step": gemm_ms / loop_ms if loop_ms > 0 else None,
        "family_ms_share": {k: (v["ms"] / loop_ms if loop_ms > 0 else None) for k, v in prof.items()},
        "note": f"precision mode {args.precision}; algorithmic FLOPs (one multiply-add per weight per row; the extra passes of "
                f"the 3xTF32 split are not counted); measured with one CUDA event per launch over {prof_steps} un-graphed loop "
                "iterations at the bench batch",
    }

    line = {
        "metric": "clips/sec", "value": value, "unit": "clips/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": {"fp32": "f32", "tf32x3": "f32 (3xTF32 tensor-core split, fp32 accumulate)", "tf32": "tf32"}[args.precision],
        "data": "synthetic", "denoise_clip_steps_per_s": value * NUM_STEPS,
        "config": workload_config(args, world),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "clips/s", "h2d_bytes_per_step": int(wave_host.numel() * 4) * world,
                "d2h_bytes_per_step": int(gB * T * 32 * 4) * world},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "algorithmic_tflops_whole_step": 2 * gB * fl["total"] * NUM_STEPS / (ms_step / 1000.0) / 1e12,
    }

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
                                  "ms_per_denoise_step": ms1 / 3 / NUM_STEPS, "ms_per_clip": ms1 / 3}
        if not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            te, ti = cpu_sample(sd, args.cpu_iters, threads)
            line["cpu_baseline"] = {
                "value": cpu_clips_per_s(te, ti), "unit": "clips/s", "cores": threads, "kind": "port",
                "sample": f"oracle (torch-CPU restatement with the reference's cost structure) on {threads} host threads: encoder "
                          f"for one 5 s clip ({te:.2f} s) + {args.cpu_iters} of {NUM_STEPS} loop iterations at batch 1 "
                          f"({ti * 1000:.1f} ms each), extrapolated to the full loop",
            }
    if rank == 0:
        out.write(json.dumps(line) + "\n")
        out.flush()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
