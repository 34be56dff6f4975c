"""Attribute the tcgen05 GEMM's time to its parts by disabling them (said_op_gemm_tc_bench)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from said_b200._lib import Engine

eng = Engine(torch.device("cuda:0"))
M = 38400
print("M=%d N=192; ms per launch; dbg bits: 2 no-W 4 no-epilogue-IO 8 no-MMA" % M)
for K in (192, 576, 768):
    for ns in (3, 1):
        row = []
        for dbg in (0, 2, 4, 8, 2 | 4, 2 | 4 | 8):
            row.append("%d:%.3f" % (dbg, eng.op_gemm_tc_bench(M, K, ns, True, dbg, 10)))
        flops = 2.0 * M * 192 * K
        print("K=%4d nsplit=%d  " % (K, ns), "  ".join(row), "  | full = %.1f TFLOP/s" % (flops / (float(row[0].split(':')[1]) * 1e-3) / 1e12))
