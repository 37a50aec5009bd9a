"""Where the host-buffer (e2e) step of bench.py spends its time: the device-resident step, then the streamed step with the
staging path and with the page-locked mirror (engine option pinned_mirror), each with the host time of its two calls.

    python tools/e2e_probe.py [--steps 300]
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "amazon-dsstne_b200"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=300)
    a = ap.parse_args()
    import torch
    import bench
    import dsstne_b200
    from dsstne_b200 import datagen, engine
    wl = bench.workload("c2")
    engine.startup(0, 1, 0, None, seed=12134)
    stream = torch.cuda.Stream(0)
    torch.cuda.set_stream(stream)
    engine.set_stream(stream.cuda_stream)
    engine.set_option("fuse_output_gemm", 1)
    engine.set_option("pdl", 1)
    data = bench.make_data(wl, 16)
    B = wl["batch"]
    H = bench.HYPER
    n_batches = data.examples // B
    batches = []
    for b in range(n_batches):
        s0 = int(data.start[b * B])
        st = (data.start[b * B:(b + 1) * B] - np.uint64(s0)).astype(np.uint64)
        en = (data.end[b * B:(b + 1) * B] - np.uint64(s0)).astype(np.uint64)
        ix = np.ascontiguousarray(data.index[s0:int(data.end[(b + 1) * B - 1])])
        batches.append((st, en, ix))
    big = max(batches, key=lambda b: len(b[2]))
    first = datagen.HostCsr(big[0], big[1], big[2], wl["items"])

    def run(mirror, resident):
        engine.set_option("pinned_mirror", mirror)
        if resident:
            ds_in = engine.Dataset.from_host_csr("gl_input", data)
            ds_out = engine.Dataset.from_host_csr("gl_output", data)
        else:
            ds_in = engine.Dataset("gl_input", first.start, first.end, first.index, wl["items"])
            ds_out = engine.Dataset("gl_output", first.start, first.end, first.index, wl["items"])
        net = engine.Network(engine.autoencoder_json(wl["hidden"], smce=bench.SMCE, init=("Gaussian", 0.01, 0.0)), B, [ds_in, ds_out])
        net.set_training_mode(dsstne_b200.SGD)
        net.set_gemm_mode(2)
        t_load = t_step = 0.0

        def step(i, timed=False):
            nonlocal t_load, t_step
            t0 = time.perf_counter()
            pos = 0
            if resident:
                pos = (i % n_batches) * B
            else:
                st, en, ix = batches[i % n_batches]
                ds_in.load_sparse(st, en, ix)
                ds_out.load_sparse(st, en, ix)
            t1 = time.perf_counter()
            net.train_step(pos, H["alpha"], H["lam"], H["lam1"], H["mu"], H["mu1"])
            t2 = time.perf_counter()
            if timed:
                t_load += t1 - t0
                t_step += t2 - t1
        for i in range(20):
            step(i)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(a.steps):
            step(i, True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        # second pass with the engine's step trace on (its events cost a few microseconds per step: not part of the timing above)
        import ctypes as C
        engine.set_option("step_trace", 1)
        for i in range(a.steps):
            step(i)
        out = (C.c_double * 9)()
        n = dsstne_b200.lib().dsb200_engine_step_trace(out, 9)
        engine.set_option("step_trace", 0)
        trace = [round(x, 1) for x in out[:n]]
        net.close()
        return dt / a.steps * 1e3, t_load / a.steps * 1e3, t_step / a.steps * 1e3, trace

    for name, mirror, resident in (("device-resident", 0, True), ("streamed, staging", 0, False), ("streamed, pinned mirror", 1, False)):
        ms, load, stp, trace = run(mirror, resident)
        print(f"{name:34s} {ms:.4f} ms / step   host: load_sparse x2 {load:.4f} ms, train_step {stp:.4f} ms   {B / ms * 1e3 / 1e6:.3f} M samples/s", flush=True)
        print(f"    host us [prep, forward, loss, backward launches; wait for loss; update launches] {trace[:6]}   device us [start->loss, loss->end, end->next start] {trace[6:]}", flush=True)


if __name__ == "__main__":
    main()
