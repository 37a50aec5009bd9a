"""Times the output layer's forward of BASELINE config 2 both ways (run on the GPU box): dsb200_gemm_fwd_bias_act + dsb200_output_pass
against dsb200_gemm_fwd_output_pass (experimental: the output pass in the GEMM epilogue).  CUDA events, 20 launches each."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import dsstne_b200 as dsb
from helpers import ml20m, to_device


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def main():
    ctx = dsb.Context(0)
    B, k, n = 1024, 128, 27278
    g = torch.Generator(device="cuda").manual_seed(3)
    A = torch.rand(B, k, device="cuda", generator=g)
    W = torch.randn(k, n, device="cuda", generator=g) * 0.1
    bias = torch.randn(n, device="cuda", generator=g) * 0.5 - 2.0
    ds = to_device(dsb, ml20m(examples=B, width=n))
    z, delta = torch.empty(B, n, device="cuda"), torch.empty(B, n, device="cuda")
    acc = torch.zeros(1, dtype=torch.int64, device="cuda")
    ctx.set_params(smce=(1.0, 0.0, 1.0, 1.0))
    ctx.set_option("gemm_mode", 2)

    def two_calls():
        ctx.gemm_fwd_bias_act(A, W, bias, 3, z)
        ctx.output_pass(ds, 3, dsb.ACT_SIGMOID, 0, B, z, None, delta, acc)

    def fused():
        ctx.gemm_fwd_output_pass(ds, 3, dsb.ACT_SIGMOID, 0, A, W, bias, None, delta, acc)

    print(f"gemm_fwd_bias_act + output_pass: {timed(two_calls):.1f} us", flush=True)
    print(f"gemm_fwd_output_pass (fused):    {timed(fused):.1f} us", flush=True)


if __name__ == "__main__":
    main()
