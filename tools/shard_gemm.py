import os, sys, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tools")
import dsstne_b200 as dsb
from stream_bench import timed
ctx = dsb.Context(0)
g = torch.Generator(device="cuda").manual_seed(3)
for (B, k, n) in ((1024, 128, 3410), (1024, 128, 6820), (1024, 128, 13639)):
    A = torch.rand(B, k, device="cuda", generator=g); W = torch.randn(k, n, device="cuda", generator=g) * 0.1
    D = torch.randn(B, n, device="cuda", generator=g) * 0.01
    G = torch.empty(k, n, device="cuda"); Dp = torch.empty(B, k, device="cuda")
    ref_dw = (A.double().t() @ D.double()) * (-1.0 / B); ref_dx = D.double() @ W.double().t()
    ctx.set_option("gemm_mode", 2)
    for mw in (2048, 1024):
        ctx.set_option("gemm_tc_min_work", mw)
        ctx.gemm_dw(A, D, G, -1.0 / B); ctx.gemm_dx(D, W, Dp); ctx.sync()
        e1 = ((G.double() - ref_dw).abs().max() / ref_dw.abs().max()).item(); e2 = ((Dp.double() - ref_dx).abs().max() / ref_dx.abs().max()).item()
        print(f"B={B} k={k} n={n} min_work {mw}: dw {timed(lambda: ctx.gemm_dw(A, D, G, -1.0 / B)):6.1f} us (err {e1:.1e})   dx {timed(lambda: ctx.gemm_dx(D, W, Dp)):6.1f} us (err {e2:.1e})", flush=True)
