"""Where the fp16x3 GEMM's time goes: each production shape timed with parts of the kernel disabled
(said_op_gemm_h_bench dbg bits: 1 no activation TMA loads, 2 no weight copies, 4 no epilogue I/O, 8 no MMAs).

    python profiles/gemm_h_breakdown.py > gpurun_out/r2_gemm_h_breakdown.md
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from said_b200._lib import Engine  # noqa: E402

eng = Engine(torch.device("cuda:0"))
M = 38528
shapes = [("conv K=576", 192, 3, 192, 1), ("conv K=1152", 384, 3, 192, 1), ("plain K=192 N=192", 192, 1, 192, 1), ("qkv K=192 N=576 (no residual)", 192, 1, 576, 0),
          ("GEGLU K=192 N=1536 (GEGLU epilogue, pair output)", 192, 1, 1536, 2), ("ffp K=960", 960, 1, 192, 1),
          ("fused feed-forward (GEGLU + ffp in one kernel)", 192, 1, 192, 3)]
modes = [("full", 0), ("no A loads", 1), ("no W copies", 2), ("no A, no W", 3), ("no epilogue I/O", 4), ("no MMAs", 8), ("MMAs only", 7), ("nothing", 15)]
print("| shape (M = 38528) | " + " | ".join(n for n, _ in modes) + " |")
print("|---|" + "---:|" * len(modes))
for name, cin, taps, n, res in shapes:
    cells = []
    for _, dbg in modes:
        cells.append(f"{1000 * eng.op_gemm_h_bench(M, cin, taps, n, res, dbg, 20):.1f}")
    print(f"| {name} | " + " | ".join(cells) + " |  (us)")
