import os, sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
import dsstne_b200 as dsb
ctx = dsb.Context(0)
B, N, K = 2048, 1000000, 100
key = torch.empty(B, K, device="cuda"); val = torch.empty(B, K, dtype=torch.int32, device="cuda")
def t(scores, label):
    for _ in range(2): ctx.topk(scores, K, key, val)
    ctx.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): ctx.topk(scores, K, key, val)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{label}: {ms:.3f} ms  {4*B*N/ms/1e6:.0f} GB/s", flush=True)
g = torch.Generator(device="cuda").manual_seed(1)
s = torch.rand(B, N, device="cuda", generator=g)
t(s, "uniform random")
s2 = torch.full((B, N), -float("inf"), device="cuda")
t(s2, "all -inf (nothing ever appended)")
s3 = s.clone(); s3[:, :4096] = 2.0    # the first select sets the threshold above everything that follows
t(s3, "first 4096 large (one select, then nothing appended)")
c = torch.empty(B * N, device="cuda"); 
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): c.copy_(s.view(-1))
e1.record(); torch.cuda.synchronize()
print(f"torch copy: {e0.elapsed_time(e1)/5:.3f} ms  {8*B*N/(e0.elapsed_time(e1)/5)/1e6:.0f} GB/s (read+write)")
e0.record()
for _ in range(5): m = s.max()
e1.record(); torch.cuda.synchronize()
print(f"torch max: {e0.elapsed_time(e1)/5:.3f} ms  {4*B*N/(e0.elapsed_time(e1)/5)/1e6:.0f} GB/s (read)")
