"""Times the small-dense-layer kernels of csrc/dense_small.cu on the hidden-layer shape of BASELINE config 2 (run on the GPU box)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import dsstne_b200 as dsb
from stream_bench import timed

ctx = dsb.Context(0)
B, k, n = 1024, 128, 128
if len(sys.argv) > 3:
    B, k, n = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
g = torch.Generator(device="cuda").manual_seed(3)
A = torch.rand(B, k, device="cuda", generator=g)
W = torch.randn(k, n, device="cuda", generator=g) * 0.1
bias = torch.randn(n, device="cuda", generator=g) * 0.1
D = torch.randn(B, n, device="cuda", generator=g) * 0.01
Cm = torch.empty(B, n, device="cuda")
Dp = torch.empty(B, k, device="cuda")
G = torch.empty(k, n, device="cuda")
ctx.set_option("gemm_mode", 2)
print(f"B={B} k={k} n={n}")
print(f"gemm_fwd_bias_act : {timed(lambda: ctx.gemm_fwd_bias_act(A, W, bias, dsb.ACT_SIGMOID, Cm), 50):6.1f} us")
print(f"gemm_dx_hadamard  : {timed(lambda: ctx.gemm_dx_hadamard(D, W, dsb.ACT_SIGMOID, A, Dp), 50):6.1f} us")
print(f"dense_update (SGD): {timed(lambda: ctx.dense_update(dsb.SGD, -1.0 / B, A, D, 0.01, 0.0, 0.0, 0.0, 0.0, 0.0, None, None, W, None, None, bias), 50):6.1f} us")
ctx.set_option("gemm_mode", 0)
print(f"gemm_fwd (fp32)   : {timed(lambda: ctx.gemm_fwd(A, W, Cm, 0.0), 50):6.1f} us")
print(f"gemm_dw  (fp32)   : {timed(lambda: ctx.gemm_dw(A, D, G, 1.0), 50):6.1f} us")
print(f"gemm_dx  (fp32)   : {timed(lambda: ctx.gemm_dx(D, W, Dp), 50):6.1f} us")
