"""Bring-up timing of the fused output-layer forward (csrc/gemm_stream.cu out_fwd_kernel) under the gemm_debug switches (run on the GPU box)."""
import os
import sys

import ctypes as C

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import dsstne_b200 as dsb
from helpers import ml20m, to_device
from stream_bench import timed

ctx = dsb.Context(0)
B, k, n = 1024, 128, 27278
if len(sys.argv) > 3:
    B, k, n = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
g = torch.Generator(device="cuda").manual_seed(3)
A = torch.rand(B, k, device="cuda", generator=g)
W = torch.randn(k, n, device="cuda", generator=g) * 0.1
bias = torch.randn(n, device="cuda", generator=g) * 0.5 - 2.0
ds = to_device(dsb, ml20m(examples=B, width=n))
delta = torch.empty(B, n, device="cuda")
parts = torch.empty(4 * ((B + 127) // 128), n, device="cuda")
acc = torch.zeros(1, dtype=torch.int64, device="cuda")
ctx.set_params(smce=(1.0, 0.0, 1.0, 1.0))
for mode in (2,):
    ctx.set_option("gemm_mode", mode)
    for dbg, what in ((0, "full"), (8192, "no stores"), (2048, "no element math, no stores"), (1024, "no MMAs"), (1024 + 8192, "no MMAs, no stores"), (1024 + 2048, "neither"), (1024 + 2048 + 16384, "neither, no TMA"), (1024 + 2048 + 16384 + 32768, "neither, no TMA, no W loads"), (16384, "full but no TMA"), (2048 + 16384, "MMA only, no TMA")):
        ctx.set_option("gemm_debug", dbg)
        t = timed(lambda: ctx.gemm_fwd_output_pass(ds, 3, dsb.ACT_SIGMOID, 0, A, W, bias, None, delta, acc, parts))
        print(f"mode {mode} {what:32s} {t:7.1f} us", flush=True)
ctx.set_option("gemm_debug", 0)
ctx.set_option("profile", 1)
for _ in range(20):
    ctx.gemm_fwd_output_pass(ds, 3, dsb.ACT_SIGMOID, 0, A, W, bias, None, delta, acc, parts)
buf = C.create_string_buffer(1 << 14)
dsb.lib().dsb200_profile_report(ctx.h, buf, C.c_size_t(len(buf)))
print(buf.value.decode())

# cycle counters (gemm_debug & 65536): per CTA [0..4] worker warp 0, [5..9] worker warp 15: total, wait accFull, tcgen05.ld + arrive,
# element math + stores, W slice load / store;  [10..14] MMA thread: total, wait W slice, wait drained accumulator, wait B (TMA), issue
import numpy as np
for dbg, what in ((65536, "full"), (65536 + 2048, "MMA only"), (65536 + 1024, "no MMAs")):
    ctx.set_option("gemm_debug", dbg)
    for _ in range(3):
        ctx.gemm_fwd_output_pass(ds, 3, dsb.ACT_SIGMOID, 0, A, W, bias, None, delta, acc, parts)
    out = np.zeros(256 * 16, dtype=np.uint64)
    dsb.lib().dsb200_debug_counters(ctx.h, out.ctypes.data_as(C.c_void_p), C.c_size_t(out.size))
    c = out.reshape(256, 16)[:148].astype(np.float64) / 1e3
    names = ["w0 total", "w0 wait accFull", "w0 ld+arrive", "w0 math+stores", "w0 W slice", "w15 total", "w15 wait accFull", "w15 ld+arrive", "w15 math+stores",
             "w15 W slice", "mma total", "mma wait W", "mma wait accEmpty", "mma wait B", "mma issue"]
    print(what, "(k-cycles, mean / max over CTAs)")
    for i, nm in enumerate(names):
        print(f"    {nm:20s} {c[:, i].mean():8.1f} {c[:, i].max():8.1f}")
ctx.set_option("gemm_debug", 0)
