"""Times the three output-layer GEMMs of BASELINE config 2 on the kernels of csrc/gemm_stream.cu against the general tcgen05 kernel
(csrc/gemm_tc.cu) and the unfused forward (run on the GPU box).  CUDA events, 20 launches each, L2 not flushed (operands 112 MB)."""
import os
import sys

import ctypes as C
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
    if len(sys.argv) > 3:
        B, k, n = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    g = torch.Generator(device="cuda").manual_seed(3)
    A = torch.rand(B, k, device="cuda", generator=g)
    W = torch.randn(k, n, device="cuda", generator=g) * 0.1
    bias = torch.randn(n, device="cuda", generator=g) * 0.5 - 2.0
    ds = to_device(dsb, ml20m(examples=B, width=n))
    z, delta = torch.empty(B, n, device="cuda"), torch.empty(B, n, device="cuda")
    G = torch.empty(k, n, device="cuda")
    Dp = torch.empty(B, k, device="cuda")
    parts = torch.empty(4 * ((B + 127) // 128), n, device="cuda")
    acc = torch.zeros(1, dtype=torch.int64, device="cuda")
    ctx.set_params(smce=(1.0, 0.0, 1.0, 1.0))
    ctx.set_option("gemm_mode", 2)

    def two_calls():
        ctx.gemm_fwd_bias_act(A, W, bias, 3, z)
        ctx.output_pass(ds, 3, dsb.ACT_SIGMOID, 0, B, z, None, delta, acc)

    def fused():
        ctx.gemm_fwd_output_pass(ds, 3, dsb.ACT_SIGMOID, 0, A, W, bias, None, delta, acc, parts)

    print(f"shape B={B} k={k} n={n}")
    print(f"gemm_fwd_bias_act + output_pass: {timed(two_calls):.1f} us", flush=True)
    print(f"gemm_fwd_output_pass (fused):    {timed(fused):.1f} us", flush=True)
    for stream in (0, 1):
        ctx.set_option("gemm_stream", stream)
        print(f"gemm_stream={stream}: dw {timed(lambda: ctx.gemm_dw(A, delta, G, -1.0 / B)):.1f} us   dx {timed(lambda: ctx.gemm_dx(delta, W, Dp)):.1f} us", flush=True)
    for mode, dbg, what in ((1, 0, "1xTF32"), (1, 512, "1xTF32, no A loads"), (2, 512, "3xTF32, no A loads"), (2, 1024, "3xTF32, no MMA")):
        ctx.set_option("gemm_mode", mode); ctx.set_option("gemm_debug", dbg)
        print(f"{what}: dw {timed(lambda: ctx.gemm_dw(A, delta, G, -1.0 / B)):.1f} us   dx {timed(lambda: ctx.gemm_dx(delta, W, Dp)):.1f} us", flush=True)
    ctx.set_option("gemm_mode", 2); ctx.set_option("gemm_debug", 0)
    ctx.set_option("profile", 1)
    for _ in range(20):
        fused(); ctx.gemm_dw(A, delta, G, -1.0 / B); ctx.gemm_dx(delta, W, Dp)
    buf = C.create_string_buffer(1 << 14)
    dsb.lib().dsb200_profile_report(ctx.h, buf, C.c_size_t(len(buf)))
    print(buf.value.decode())


if __name__ == "__main__":
    main()
