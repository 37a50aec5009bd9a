"""Bring-up timing sweep of the tcgen05 GEMM: ring depth and the gemm_debug switches (1 no loads, 2 no MMA, 4 no stores, 8 no lo)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dsstne_b200 as dsb
ctx = dsb.Context(0)
B, k, n = 1024, 128, 27278
g = torch.Generator(device="cuda").manual_seed(1)
A = torch.randn(B, k, device="cuda", generator=g); W = torch.randn(k, n, device="cuda", generator=g) * 0.1
D = torch.randn(B, n, device="cuda", generator=g) * 0.1
C = torch.zeros(B, n, device="cuda"); G = torch.zeros(k, n, device="cuda"); Dp = torch.zeros(B, k, device="cuda")
ops = (("fwd", lambda: ctx.gemm_fwd(A, W, C, beta=0.0)), ("dw", lambda: ctx.gemm_dw(A, D, G, -1.0 / B)), ("dx", lambda: ctx.gemm_dx(D, W, Dp)))
def t(fn):
    for _ in range(2): fn()
    ctx.sync(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 10 * 1e3
for mode in (2, 1):
    ctx.set_option("gemm_mode", mode)
    for splits in (0,):
        ctx.set_option("gemm_splits", splits)
        for dbg in (0, 2, 4, 8, 16, 6, 14, 30, 31):
            ctx.set_option("gemm_debug", dbg)
            print(f"mode={mode} splits={splits} debug={dbg}: " + " ".join(f"{name} {t(fn):7.1f}" for name, fn in ops), flush=True)
