"""Bring-up probe for the tcgen05 GEMM (run on the GPU box): errors of every mode / descriptor convention against float64,
and timing against the cuBLAS fp32 path.  Not a test -- tests/test_gpu_gemm.py is."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dsstne_b200 as dsb


def rel(a, b):
    b = b.double()
    return float(((a.double() - b).abs() / (b.abs() + b.pow(2).mean().sqrt())).max())


def main():
    ctx = dsb.Context(0)
    g = torch.Generator(device="cuda").manual_seed(1)
    shapes = [(256, 128, 256), (1024, 128, 128), (1024, 128, 27278), (300, 70, 1000), (130, 33, 259)]
    for swap in (0,):
        for mode, name in ((2, "tf32x3"), (1, "tf32")):
            ctx.set_option("gemm_mode", mode)
            for (B, k, n) in shapes:
                A = torch.randn(B, k, device="cuda", generator=g)
                W = torch.randn(k, n, device="cuda", generator=g) * 0.1
                D = torch.randn(B, n, device="cuda", generator=g) * 0.1
                C = torch.zeros(B, n, device="cuda")
                ctx.gemm_fwd(A, W, C, beta=0.0)
                G = torch.zeros(k, n, device="cuda")
                ctx.gemm_dw(A, D, G, -1.0 / B)
                Dp = torch.zeros(B, k, device="cuda")
                ctx.gemm_dx(D, W, Dp)
                ctx.sync()
                e1 = rel(C, A.double() @ W.double())
                e2 = rel(G, (-1.0 / B) * (A.double().T @ D.double()))
                e3 = rel(Dp, D.double() @ W.double().T)
                print(f"swap={swap} {name} B={B} k={k} n={n}: fwd {e1:.2e} dw {e2:.2e} dx {e3:.2e}", flush=True)
    B, k, n = 1024, 128, 27278
    A = torch.randn(B, k, device="cuda", generator=g)
    W = torch.randn(k, n, device="cuda", generator=g) * 0.1
    D = torch.randn(B, n, device="cuda", generator=g) * 0.1
    C = torch.zeros(B, n, device="cuda")
    G = torch.zeros(k, n, device="cuda")
    Dp = torch.zeros(B, k, device="cuda")
    bias = torch.randn(n, device="cuda", generator=g)
    for mode, loader, name in ((0, -1, "fp32-cublas"), (2, -1, "tf32x3-auto"), (2, 1, "tf32x3-regload"), (2, 0, "tf32x3-cpasync"), (1, -1, "tf32-auto")):
        ctx.set_option("gemm_mode", mode)
        ctx.set_option("gemm_loader", loader)
        for fn, label in ((lambda: ctx.gemm_fwd(A, W, C, beta=0.0), "fwd"), (lambda: ctx.gemm_dw(A, D, G, -1.0 / B), "dw"),
                          (lambda: ctx.gemm_dx(D, W, Dp), "dx"), (lambda: ctx.gemm_fwd_bias_act(A, W, bias, 0, C), "fwd+bias+sigmoid")):
            for _ in range(3):
                fn()
            ctx.sync()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            ev0.record()
            for _ in range(20):
                fn()
            ev1.record()
            torch.cuda.synchronize()
            print(f"{name} {label}: {ev0.elapsed_time(ev1) / 20 * 1e3:.1f} us", flush=True)


if __name__ == "__main__":
    main()
