import torch, time, sys
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from said_b200.model.diffusion import SAID_UNet1D
from said_b200.synth import synthetic_batch, synthetic_state_dict
m = SAID_UNet1D(); m.load_state_dict(synthetic_state_dict(0)); m.to("cuda:0").eval()
w = synthetic_batch(1, 5.0).to("cuda:0")
noise = torch.randn(1, 300, 32, device="cuda:0")
for mode, rows in (("fp32", 0), ("tf32x3", 2048), ("tf32x3", 1)):
    m.precision = mode; m.tc_min_rows = rows
    for _ in range(2):
        m._run(w, noise, None, None, 1000, 0.2, 2.0, 0.0, 0.0, 300, False, False, None)
    torch.cuda.synchronize(); t0 = time.time()
    m._run(w, noise, None, None, 1000, 1.0, 2.0, 0.0, 0.0, 300, False, False, None)
    torch.cuda.synchronize(); print(mode, rows, "B=1 1000 steps: %.1f ms" % ((time.time() - t0) * 1e3))
w8 = synthetic_batch(8, 5.0).to("cuda:0"); n8 = torch.randn(8, 300, 32, device="cuda:0")
for mode, rows in (("fp32", 0), ("tf32x3", 2048), ("tf32x3", 1)):
    m.precision = mode; m.tc_min_rows = rows
    m._run(w8, n8, None, None, 1000, 0.1, 2.0, 0.0, 0.0, 300, False, False, None)
    torch.cuda.synchronize(); t0 = time.time()
    m._run(w8, n8, None, None, 1000, 0.5, 2.0, 0.0, 0.0, 300, False, False, None)
    torch.cuda.synchronize(); print(mode, rows, "B=8 500 steps: %.1f ms" % ((time.time() - t0) * 1e3))
