"""Fused feed-forward (ffn_h.cuh) diagnostics: time with parts disabled (dbg bits 1 no activation loads, 2 no weight copies,
4 no epilogue work, 8 no MMAs) and, with bit 16, clock64 stamps of one tile (printed by the library on stderr)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from said_b200._lib import Engine  # noqa: E402

eng = Engine(torch.device("cuda:0"))
M = 38528
modes = [int(a) for a in sys.argv[1:]] or [0, 2, 8, 10, 12, 14, 6, 7, 4, 15]
for dbg in modes:
    print(dbg, f"{1000 * eng.op_gemm_h_bench(M, 192, 1, 192, 3, dbg, 20):.1f}", flush=True)
