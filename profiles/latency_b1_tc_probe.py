import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from said_b200.model.diffusion import SAID_UNet1D
from said_b200.synth import synthetic_batch, synthetic_state_dict
m = SAID_UNet1D(); m.load_state_dict(synthetic_state_dict(0)); m.to("cuda:0").eval()
wave = synthetic_batch(1, 5.0).to("cuda:0")
torch.manual_seed(0); noise = torch.randn(1, 300, 32, device="cuda:0")
def run():
    with torch.no_grad():
        return m._run(wave, noise, None, None, 1000, 1.0, 2.0, 0.0, 0.0, 300, False, False, None)
res = {}
for rows in (2048, 256):
    m.tc_min_rows = rows
    for _ in range(2): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): out = run()
    e1.record(); torch.cuda.synchronize()
    res[rows] = out.result.clone()
    print("tc_min_rows", rows, "ms/step", e0.elapsed_time(e1) / 3 / 1000)
print("max diff between modes", float((res[2048] - res[256]).abs().max()))
