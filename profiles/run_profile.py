"""Short un-graphed run of the hot path for ncu: encoder + K/V hoist + a few loop iterations.

    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
        python profiles/run_profile.py --batch 64 --iters 3
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from said_b200.model.diffusion import SAID_UNet1D  # noqa: E402
from said_b200.synth import synthetic_batch, synthetic_state_dict  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--seconds", type=float, default=5.0)
ap.add_argument("--graph", action="store_true")
ap.add_argument("--precision", default=None)
ap.add_argument("--encoder", default=None, help="encoder precision (fp32 keeps the fp16x3 GEMM kernels out of the once-per-batch part)")
a = ap.parse_args()
m = SAID_UNet1D()
m.load_state_dict(synthetic_state_dict(0))
m.to("cuda:0").eval()
m.use_cuda_graph = a.graph
if a.precision:
    m.precision = a.precision
if a.encoder:
    m.encoder_precision = a.encoder
wave = synthetic_batch(a.batch, a.seconds).to("cuda:0")
T = int(wave.shape[1] / 16000 * 60)
torch.manual_seed(0)
noise = torch.randn(a.batch, T, 32, device="cuda:0")
with torch.no_grad():
    out = m._run(wave, noise, None, None, 1000, a.iters / 1000.0, 2.0, 0.0, 0.0, T, False, False, None)
torch.cuda.synchronize()
print("ok", float(out.result.mean()))
