"""Scaling of the tcgen05 row GEMM with the number of rows (fixed per-launch cost vs steady-state rate)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pcnerf_b200 import ops

dev = torch.device("cuda:0")
B = (torch.randn(256, 256, device=dev) * 0.1).half()
bias = torch.randn(256, device=dev)
for rows in (16384, 65536, 131072, 262144, 524288, 1048576, 2097152):
    A = torch.randn(rows, 256, device=dev).half()
    for _ in range(3):
        ops.tc_rowgemm(0, A, B, None, bias)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    e0.record()
    for _ in range(n):
        ops.tc_rowgemm(0, A, B, None, bias)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / n * 1e3
    print("rows %8d  %8.1f us  %6.1f GB/s  %6.1f TFLOP/s  (%.2f us per 128-row tile per CTA pair)"
          % (rows, us, rows * 1024 / us / 1e3, 2.0 * rows * 256 * 256 / us / 1e6, us / (rows / 128 / 74)))
