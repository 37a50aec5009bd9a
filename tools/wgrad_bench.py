"""Times the sparse weight gradient + SGD (dsb200_sparse_wgrad_update) on BASELINE config 2's input layer (27,278 x 128, batch 1,024,
ML-20M-shaped CSR) and on a config-4 shard (125,000 x 1,024), unified kernel against round 1's pair.  CUDA events, 20 launches."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import dsstne_b200 as dsb
from helpers import ml20m, to_device
from oracle import oracle as orc


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


def case(ctx, N, S, B, mean):
    h = ml20m(examples=B, width=N, mean=mean)
    oc = orc.Csr(h.start, h.end, h.index)
    tstart, cap = orc.transposed_capacity(oc, N, B)
    d_start = torch.from_numpy(tstart.view(np.int32).copy()).cuda()
    d_end = torch.zeros(N, dtype=torch.int32, device="cuda")
    d_idx = torch.zeros(max(cap, 1), dtype=torch.int32, device="cuda")
    ctx.set_option("transpose_sort", 0)
    ctx.sparse_transpose(to_device(dsb, h), 0, B, N, d_start, d_end, d_idx, None, False)
    delta = torch.randn(B, S, device="cuda") * 0.01
    W = torch.randn(N, S, device="cuda") * 0.01
    ctx.set_option("wgrad_max_entries", h.nnz + 64)
    alg = 4 * S * h.nnz + 2 * 4 * S * N + 4 * h.nnz + 8 * N
    for two in (0, 1):
        ctx.set_option("wgrad_two_kernel", two)
        us = timed(lambda: ctx.sparse_wgrad_update(dsb.SGD, -1.0 / B, d_start, d_end, d_idx, None, delta, 0.025, 1e-4, 0.0, 0.0, 0.0, 0.0, None, None, W))
        print(f"N={N} S={S} nnz={h.nnz}  {'two-kernel' if two else 'unified   '}: {us:7.1f} us   {alg / us / 1e3:7.1f} GB/s algorithmic ({alg / 1e6:.1f} MB)", flush=True)
    ctx.set_option("wgrad_two_kernel", 0)


def main():
    ctx = dsb.Context(0)
    case(ctx, 27278, 128, 1024, 144.4)
    case(ctx, 125000, 1024, 1024, 18.0)          # a config-4 shard: 1/8 of the columns, 1/8 of the entries


if __name__ == "__main__":
    main()
