"""Bring-up timing matrix for the tcgen05 GEMM (run on the GPU box): the three output-layer GEMMs of BASELINE config 2
under every operand loader and the kernel's bring-up switches (loaders: 0 cp.async, 1 registers, 2 tensor-memory A, 3 coalesced tensor-memory A -- pass the ones to time as
arguments; option "gemm_debug": 1 no proxy fence, 4 no MMA, 8 no lo
pass, 16 no stores, 32 no global loads).  Results of the switched runs are wrong by construction -- only the time matters:
it tells which stage of the pipeline paces a k-iteration.  Not a test."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dsstne_b200 as dsb


def main():
    ctx = dsb.Context(0)
    g = torch.Generator(device="cuda").manual_seed(1)
    B, k, n = 1024, 128, 27278
    A = torch.randn(B, k, device="cuda", generator=g)
    W = torch.randn(k, n, device="cuda", generator=g) * 0.1
    D = torch.randn(B, n, device="cuda", generator=g) * 0.1
    C = torch.zeros(B, n, device="cuda")
    G = torch.zeros(k, n, device="cuda")
    Dp = torch.zeros(B, k, device="cuda")
    fns = (("fwd", lambda: ctx.gemm_fwd(A, W, C, beta=0.0)), ("dw", lambda: ctx.gemm_dw(A, D, G, -1.0 / B)), ("dx", lambda: ctx.gemm_dx(D, W, Dp)))
    ctx.set_option("gemm_mode", 2)
    print("loader debug " + " ".join(f"{l:>8s}" for l, _ in fns), flush=True)
    for loader in ([int(x) for x in sys.argv[1:]] or [0, 1, 2]):
        for debug in (0, 4, 16, 32, 36, 52):
            if loader == 0 and debug >= 32:
                continue
            ctx.set_option("gemm_loader", loader)
            ctx.set_option("gemm_debug", debug)
            out = []
            for _, fn in fns:
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
                out.append(ev0.elapsed_time(ev1) / 20 * 1e3)
            print(f"{loader:6d} {debug:5d} " + " ".join(f"{t:8.1f}" for t in out), flush=True)
    ctx.set_option("gemm_debug", 0)
    ctx.set_option("gemm_loader", -1)


if __name__ == "__main__":
    main()
