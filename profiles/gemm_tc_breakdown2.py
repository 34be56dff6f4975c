import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from said_b200._lib import Engine
eng = Engine(torch.device("cuda:0"))
for M in (38400, 19200, 37888):
    print("M=%d N=192; ms per launch (dbg bits: 2 skip weights, 4 skip epilogue IO, 8 skip MMAs, 32 no tail slivers)" % M)
    for K in (192, 576, 768, 1152):
        for ns in (3, 1):
            row = []
            for dbg in (0, 32, 4, 4|32):
                row.append("%d:%.1f" % (dbg, 1000*eng.op_gemm_tc_bench(M, K, ns, True, dbg, 10)))
            flops = 2.0 * M * 192 * K
            print("K=%4d ns=%d " % (K, ns), " ".join(row), " | %.1f TFLOP/s" % (flops / (float(row[0].split(':')[1]) * 1e-6) / 1e12))
