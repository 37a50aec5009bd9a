"""BASELINE config 5: top-K = 100 over a 1M-item output with an exclusion filter, batch 4,096 (scaled by --batch so the score
matrix fits), through dsb200_topk.  Prints time and algorithmic GB/s (4*B*N + 8*B*K + 4*nnz_filter bytes)."""
import argparse, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dsstne_b200 as dsb
from dsstne_b200 import datagen

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=4096)
ap.add_argument("--items", type=int, default=1000000)
ap.add_argument("--k", type=int, default=100)
a = ap.parse_args()
ctx = dsb.Context(0)
g = torch.Generator(device="cuda").manual_seed(1)
scores = torch.rand(a.batch, a.items, device="cuda", generator=g)
h = datagen.make_csr(a.batch, a.items, 144.4, dist="lognormal", col="zipf", seed=3)
fs = torch.from_numpy(h.start.view(np.int64)).cuda(); fe = torch.from_numpy(h.end.view(np.int64)).cuda()
fi = torch.from_numpy(h.index.view(np.int32)).cuda()
key = torch.empty(a.batch, a.k, device="cuda"); val = torch.empty(a.batch, a.k, dtype=torch.int32, device="cuda")
def run(filt):
    ctx.topk(scores, a.k, key, val, filt=(fs, fe, fi) if filt else None)
for filt in (False, True):
    for _ in range(2): run(filt)
    ctx.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): run(filt)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    nbytes = 4 * a.batch * a.items + 8 * a.batch * a.k + (4 * h.nnz if filt else 0)
    print(f"topk batch={a.batch} items={a.items} k={a.k} filter={filt}: {ms:.3f} ms  {nbytes / ms / 1e6:.0f} GB/s algorithmic  {a.batch / ms * 1e3:.0f} rows/s", flush=True)
# spot check against torch.topk
tk = torch.topk(scores[:8], a.k, dim=1)
run(False); ctx.sync()
print("keys match torch.topk:", bool(torch.equal(tk.values, key[:8])), "indices match:", bool(torch.equal(tk.indices.int(), val[:8])))
