"""One launch each of the tcgen05 GEMM (N = 192, plain loader, residual epilogue) at K = 1152 and K = 192 for
ncu --set full --import-source on (stall reasons per source line of the producer / MMA / epilogue roles)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from said_b200._lib import Engine

eng = Engine(torch.device("cuda:0"))
for K in (1152, 192):
    print(K, eng.op_gemm_tc_bench(38400, K, 3, True, 0, 1))
