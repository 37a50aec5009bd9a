"""Runs each big C2 GEMM a few times in one mode (for ncu):  python tools/gemm_once.py [mode] [stages]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dsstne_b200 as dsb
ctx = dsb.Context(0)
ctx.set_option("gemm_mode", int(sys.argv[1]) if len(sys.argv) > 1 else 2)
if len(sys.argv) > 2:
    ctx.set_option("gemm_stages", int(sys.argv[2]))
B, k, n = 1024, 128, 27278
g = torch.Generator(device="cuda").manual_seed(1)
A = torch.randn(B, k, device="cuda", generator=g); W = torch.randn(k, n, device="cuda", generator=g) * 0.1
D = torch.randn(B, n, device="cuda", generator=g) * 0.1
C = torch.zeros(B, n, device="cuda"); G = torch.zeros(k, n, device="cuda"); Dp = torch.zeros(B, k, device="cuda")
for _ in range(2):
    ctx.gemm_fwd(A, W, C, beta=0.0); ctx.gemm_dw(A, D, G, -1.0 / B); ctx.gemm_dx(D, W, Dp)
ctx.sync()
