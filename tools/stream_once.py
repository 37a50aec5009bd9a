"""Runs the three output-layer kernels of csrc/gemm_stream.cu a few times on BASELINE config 2's shape (for ncu)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import dsstne_b200 as dsb
from helpers import ml20m, to_device
ctx = dsb.Context(0)
ctx.set_option("gemm_mode", 2)
B, k, n = 1024, 128, 27278
g = torch.Generator(device="cuda").manual_seed(1)
A = torch.rand(B, k, device="cuda", generator=g); W = torch.randn(k, n, device="cuda", generator=g) * 0.1
bias = torch.randn(n, device="cuda", generator=g) * 0.5 - 2.0
ds = to_device(dsb, ml20m(examples=B, width=n))
delta = torch.empty(B, n, device="cuda"); G = torch.zeros(k, n, device="cuda"); Dp = torch.zeros(B, k, device="cuda")
parts = torch.empty(4 * ((B + 127) // 128), n, device="cuda")
acc = torch.zeros(1, dtype=torch.int64, device="cuda")
ctx.set_params(smce=(1.0, 0.0, 1.0, 1.0))
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    ctx.gemm_fwd_output_pass(ds, 3, dsb.ACT_SIGMOID, 0, A, W, bias, None, delta, acc, parts)
    ctx.gemm_dw(A, delta, G, -1.0 / B)
    ctx.gemm_dx(delta, W, Dp)
ctx.sync()
