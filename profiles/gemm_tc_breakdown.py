"""Time the tcgen05 GEMM variants on scratch buffers (said_op_gemm_tc_bench): dbg 0 = A and B in shared memory,
256 = A through TMEM; other bits disable parts of the shared-memory kernel (2 weights, 4 epilogue I/O, 8 MMAs)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from said_b200._lib import Engine

eng = Engine(torch.device("cuda:0"))
M = 38400
print("M=%d N=192; ms per launch" % M)
for K in (192, 576, 768, 1152):
    for ns in (3, 1):
        row = []
        for dbg in (0, 256, 4, 2 | 4):
            row.append("%d:%.3f" % (dbg, eng.op_gemm_tc_bench(M, K, ns, True, dbg, 10)))
        flops = 2.0 * M * 192 * K
        print("K=%4d nsplit=%d  " % (K, ns), "  ".join(row), "  | smem-A %.1f, tmem-A %.1f TFLOP/s" % (
            flops / (float(row[0].split(':')[1]) * 1e-3) / 1e12, flops / (float(row[1].split(':')[1]) * 1e-3) / 1e12))
