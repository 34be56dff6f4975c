import sys
sys.path.insert(0, "/root/repo")
import torch
from said_b200._lib import Engine
eng = Engine(torch.device("cuda:0"))
for M in (602, 2408):
    for (cin, taps) in ((192, 1), (192, 3), (384, 3), (960, 1)):
        for N in (32, 192):
            t = 1000 * eng.op_gemm_h_bench(M, cin, taps, N, 1, 0, 50)
            print(f"M={M} K={cin*taps} N={N}: {t:.1f} us", flush=True)
