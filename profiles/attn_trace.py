import sys, time
sys.path.insert(0, "/root/repo")
import torch
from said_b200._lib import Engine
eng = Engine(torch.device("cuda:0"))
B, T = 128, 300
qkv = torch.randn(B, T, 576, device="cuda:0")
for _ in range(3):
    out = eng.op_self_attention_h(qkv, 6)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    out = eng.op_self_attention_h(qkv, 6)
torch.cuda.synchronize()
print("us per call", (time.perf_counter() - t0) / 20 * 1e6)
