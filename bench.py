#!/usr/bin/env python
"""bench.py -- headline benchmark of the dsstne_b200 hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c4]

A "step" is one minibatch of NNNetwork::Train (transposed build, sparse-Z forward, dense layers,
fused sigmoid + SMCE loss + delta, backward GEMMs, sparse gradient fused with the optimizer, bias
updates) on BASELINE.json config 2: the MovieLens-20M-shape sparse autoencoder
27,278 -> 128 -> 128 -> 128 -> 27,278, batch 1,024, synthetic CSR at ML-20M density, SGD.
N > 1 (torchrun, one rank per GPU) runs the SAME model model-parallel (config 3): layers split by
unit, NCCL reduce-scatter / all-gather per layer boundary -> strong scaling.

Prints ONE JSON line (rank 0).  `value` = samples/s with the dataset resident in HBM; `e2e` = the
same through the host-buffer API (each step uploads its CSR batch from pinned host memory with
NNDataSet::LoadSparseData and reads the loss back); `roofline` = achieved algorithmic GB/s of the
dominant hand-written kernel (CUDA events inside the timed steps) against MEASURED_PEAKS.json;
`cpu_baseline` = the OpenMP CPU oracle on a bounded sample of the same workload.
`--impl reference` times the CPU oracle only (the reference has no CPU path and its full build needs
MPI/NetCDF/jsoncpp) with all host threads.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "sparse-AE train samples/s"
UNIT = "samples/s"
HYPER = dict(alpha=0.025, lam=1e-4, lam1=0.0, mu=0.5, mu1=0.0)     # CLI defaults of U/Train.cpp:62-66
SMCE = (1.0, 0.0, 1.0, 1.0)                                        # samples/movielens/config.json


def workload(name):
    if name == "c2":
        return dict(name="c2", items=27278, hidden=[128, 128, 128], batch=1024, mean_nnz=144.4,
                    desc="BASELINE config 2: ML-20M-shape AE 27278-128-128-128-27278, batch 1024, sigmoid/SMCE(1,0,1,1), SGD, "
                         "synthetic CSR (log-normal rows mean 144.4, Zipf-Mandelbrot columns)")
    if name == "c4":
        return dict(name="c4", items=1000000, hidden=[1024, 1024, 1024], batch=1024, mean_nnz=144.4,
                    desc="BASELINE config 4: 1M-item AE 1M-1024-1024-1024-1M, batch 1024, model parallel")
    raise SystemExit(f"unknown workload {name}")


def make_data(wl, batches, seed=12134):
    from dsstne_b200 import datagen
    return datagen.make_csr(wl["batch"] * batches, wl["items"], wl["mean_nnz"], dist="lognormal", col="zipf", seed=seed)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device):
        super().__init__(daemon=True)
        self.device, self.rows, self.stop_flag = device, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for i, n in enumerate(names):
                if len(r) > 4 + i and r[4 + i].lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def algorithmic_bytes(wl, nnz_batch, P=1):
    """SURVEY.md section 8d per-launch ALGORITHMIC bytes of each hand-written kernel family (fp32).  Families that are called
    on several layers report per call site as "name@size" (csrc/profile.cu); those are resolved by alg_of() from the size."""
    B, N, S, H = wl["batch"], wl["items"] // P, wl["hidden"][0], wl["hidden"][-1]
    return {
        # gathered weight rows + Z write (bias read) + indices + start/end
        "sparse_z_bias_act": 4 * S * nnz_batch + 4 * B * S + 4 * nnz_batch + 16 * B,
        "sparse_z": 4 * S * nnz_batch + 2 * 4 * B * S + 4 * nnz_batch + 16 * B,
        # CSR read + TIndex write + End init/final
        "sparse_transpose": 4 * nnz_batch + 16 * B + 4 * nnz_batch + 8 * N,
        # gathered delta rows + W read/write (fused SGD: no dW) + TIndex + start/end
        "sparse_wgrad_update": 4 * S * nnz_batch + 2 * 4 * S * N + 4 * nnz_batch + 8 * N,
        "sparse_wgrad": 4 * S * nnz_batch + 4 * S * N + 4 * nnz_batch + 8 * N,
        # read Z, write delta + target CSR (the activations are not stored during training)
        "output_pass": 2 * 4 * B * N + 4 * nnz_batch + 16 * B,
        # forward GEMM + activation + loss + delta in one kernel: X and W read, delta written, target CSR read (Z never exists)
        "gemm_fwd_output_pass": 4 * (B * H + H * N + B * N + N) + 4 * nnz_batch + 16 * B,
        # output-layer GEMMs: operands read once + result written once (7.15 GFLOP each on c2)
        "gemm_fwd_bias_act_tc": 4 * (B * H + H * N + B * N + N),
        "gemm_dw_tc": 4 * (B * H + B * N + H * N), "gemm_dw_stream": 4 * (B * H + B * N + H * N),
        "gemm_dx_tc": 4 * (B * N + H * N + B * H), "gemm_dx_stream": 4 * (B * N + H * N + B * H),
    }


def alg_of(name, alg, batch):
    """algorithmic bytes of one profiled call site, or None (library / bookkeeping call)"""
    if name in alg:
        return alg[name]
    fam, _, tag = name.partition("@")
    if not tag:
        return None
    n = int(tag)
    if fam == "update_weights":
        return 3 * 4 * n                                  # SGD: read g, read w, write w
    if fam == "update_biases":
        return 4 * batch * n + 8 * n                      # delta column sums + bias r/w
    if fam == "update_biases_partials":
        return 8 * 4 * n + 8 * n
    if fam == "regularization_error":
        return 4 * n
    return None


def run_ours(args, wl, rank, world, local_rank):
    import torch
    import dsstne_b200
    from dsstne_b200 import engine
    dsstne_b200.lib()                                  # fails loudly when the CUDA extension is missing
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the dsstne_b200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    nccl_id = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            buf = torch.tensor(list(engine.unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(buf, 0)
        nccl_id = bytes(buf.cpu().tolist())
    engine.startup(rank, world, local_rank, nccl_id, seed=12134)
    stream = torch.cuda.Stream(local_rank)
    torch.cuda.set_stream(stream)
    engine.set_stream(stream.cuda_stream)
    engine.set_option("p2p_exchange", 1 if args.p2p else 0)
    engine.set_option("fuse_output_gemm", 1 if args.fuse_output else 0)
    engine.set_option("pdl", 1 if args.pdl else 0)
    if args.gemm_loader >= 0:
        engine.set_option("gemm_loader", args.gemm_loader)

    n_batches = 16 if wl["name"] == "c2" else 4
    data = make_data(wl, n_batches)
    B = wl["batch"]
    ds_in = engine.Dataset.from_host_csr("gl_input", data)
    ds_out = engine.Dataset.from_host_csr("gl_output", data)
    net = engine.Network(engine.autoencoder_json(wl["hidden"], smce=SMCE, init=("Gaussian", 0.01, 0.0)), B, [ds_in, ds_out])
    net.set_training_mode(dsstne_b200.SGD)
    net.set_gemm_mode(args.gemm_mode)

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def step(i):
        return net.train_step((i % n_batches) * B, HYPER["alpha"], HYPER["lam"], HYPER["lam1"], HYPER["mu"], HYPER["mu1"])

    for i in range(args.warmup):
        step(i)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = dsstne_b200.lib().dsb200_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    loss = 0.0
    for i in range(args.steps):
        loss = step(args.warmup + i)
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = dsstne_b200.lib().dsb200_launch_count() - launches0
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    sampler.stop_flag = True

    # ---- per-kernel device time inside the same steps (second pass with event pairs on) ----
    engine.set_option("profile", 1)
    pk0, pk1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    prof_steps = min(args.steps, 20)
    barrier()
    pk0.record(stream)
    for i in range(prof_steps):
        step(i)
    pk1.record(stream)
    barrier()
    prof = engine.profile_report()
    engine.set_option("profile", 0)
    prof_ms = pk0.elapsed_time(pk1)

    # ---- e2e: host buffers in, loss out, through the reference-facing API every step ----
    try:
        e2e = run_e2e(args, wl, data, engine, dsstne_b200, stream, world)
    except Exception as exc:                             # the device-resident number above must still be reported
        e2e = {"value": None, "unit": UNIT, "h2d_bytes_per_step": None, "d2h_bytes_per_step": None, "error": repr(exc)[:300]}

    result = None
    if rank == 0:
        nnz_batch = data.nnz / n_batches
        alg = algorithmic_bytes(wl, nnz_batch, world)
        peak, peak_src = peaks()
        kern = {}
        for name, (calls, tot) in prof.items():
            per = tot / max(calls, 1)
            # share: the family's device time per step over the REAL step time of the timed region (the profile pass itself is stretched
            # by its spin kernels; with work on two streams the shares can add up to more than 1)
            kern[name] = {"calls_per_step": calls / prof_steps, "ms_per_call": round(per, 5), "share": round((tot / prof_steps) / (ms / args.steps), 4)}
            ab = alg_of(name, alg, B)
            if ab is not None:
                kern[name]["algorithmic_GBs"] = round(ab / (per * 1e-3) / 1e9, 1)
                kern[name]["algorithmic_bytes"] = int(ab)
        ours = {k: v for k, v in kern.items() if "algorithmic_GBs" in v}
        dom = max(ours, key=lambda k: ours[k]["share"]) if ours else None
        roof = None
        if dom:
            a = ours[dom]["algorithmic_GBs"]
            traffic, tsrc = None, None
            tpath = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "traffic_c2.json")
            if wl["name"] == "c2" and world == 1 and os.path.exists(tpath):
                tj = json.load(open(tpath))
                traffic = tj.get(dom)                               # DRAM bytes per launch from the committed ncu capture named in "_source"
                tsrc = tj.get("_source")
            roof = {"kernel": dom, "bound": "hbm", "achieved": a, "peak": peak, "unit": "GB/s", "frac": round(a / peak, 4),
                    "traffic": traffic, "traffic_source": tsrc, "peak_source": peak_src, "share_of_step": ours[dom]["share"],
                    "algorithmic_bytes_per_launch": ours[dom]["algorithmic_bytes"],
                    "note": "3xTF32 tcgen05 kernel of the output layer: 56 flop per algorithmic byte, below the machine balance of the tf32 pipe, "
                            "so the HBM roofline is the bound that applies" if dom.startswith("gemm") else None}
        value = args.steps * B / (ms * 1e-3)
        result = {"metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                  "ms_per_step": round(ms / args.steps, 4), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                  "dtype": "f32", "data": "synthetic",
                  "config": {"workload": wl["desc"], "parallelism": "single GPU" if world == 1 else f"model-parallel mp{world} ({'peer-memory kernels' if args.p2p else 'NCCL'})",
                             "gemm": ["exact fp32 (SIMT kernel)", "tcgen05 TF32", "tcgen05 3xTF32 (fp32-grade, bound 3e-5)"][args.gemm_mode],
                             "l2": "no explicit flush: every step streams >= 3 x 112 MB of output-layer Z/delta through the 126 MB L2 "
                                   "and a different CSR batch; weights stay L2-resident exactly as in real training",
                             "last_loss": round(float(loss), 3)},
                  "clocks": sampler.summary(), "gpu_launches": int(launches), "roofline": roof, "kernels": kern}
        if e2e:
            result["e2e"] = e2e
        if args.cpu_steps > 0 and world == 1:
            result["cpu_baseline"] = cpu_baseline(wl, data, sample_steps=args.cpu_steps)
        else:                                            # --cpu-steps 0: only for the side workloads (one c4 oracle step is minutes of host time)
            result["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "skipped (--cpu-steps 0)"}
    net.close()
    ds_in.close(); ds_out.close()
    del data
    # ---- side records (BASELINE configs 4 and 5): the 1M-item layer whose scaling target BASELINE.json states, and top-K ----
    if args.side and wl["name"] == "c2":
        for key, fn in (("c4", run_side_c4), ("c5", run_side_c5)):
            try:
                rec = fn(args, engine, dsstne_b200, stream, rank, world, local_rank)
            except Exception as exc:
                rec = {"error": repr(exc)[:300]}
            if result is not None:
                result[key] = rec
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    engine.shutdown()
    return result


def _max_over_ranks(ms, world):
    if world > 1:
        import torch
        import torch.distributed as dist
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    return ms


def run_side_c4(args, engine, dsstne_b200, stream, rank, world, local_rank):
    """BASELINE config 4 (1M items, 3 x 1,024 hidden, batch 1,024; model parallel when N > 1): a short timed run after the
    headline workload so that the driver's 1/2/4/8-GPU lines carry the config the 0.85 scaling target is quoted on."""
    import torch
    wl = workload("c4")
    t0 = time.perf_counter()
    data = make_data(wl, 2)
    B = wl["batch"]
    ds_in = engine.Dataset.from_host_csr("gl_input", data)
    ds_out = engine.Dataset.from_host_csr("gl_output", data)
    net = engine.Network(engine.autoencoder_json(wl["hidden"], smce=SMCE, init=("Gaussian", 0.01, 0.0)), B, [ds_in, ds_out])
    net.set_training_mode(dsstne_b200.SGD)
    net.set_gemm_mode(args.gemm_mode)
    setup = time.perf_counter() - t0
    steps, warm = args.c4_steps, 3

    def step(i):
        return net.train_step((i % 2) * B, HYPER["alpha"], HYPER["lam"], HYPER["lam1"], HYPER["mu"], HYPER["mu1"])

    for i in range(warm):
        step(i)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for i in range(steps):
        loss = step(warm + i)
    ev1.record(stream)
    torch.cuda.synchronize()
    ms = _max_over_ranks(ev0.elapsed_time(ev1), world)
    engine.set_option("profile", 1)
    for i in range(2):
        step(i)
    prof = engine.profile_report()
    engine.set_option("profile", 0)
    net.close(); ds_in.close(); ds_out.close()
    top = sorted(((n, t / max(c, 1)) for n, (c, t) in prof.items()), key=lambda x: -x[1])[:8]
    return {"workload": wl["desc"], "metric": METRIC, "value": round(steps * B / (ms * 1e-3), 1), "unit": UNIT, "ms_per_step": round(ms / steps, 3),
            "steps": steps, "warmup": warm, "n_gpus": world, "scaling": "strong", "setup_s": round(setup, 1), "last_loss": round(float(loss), 3),
            "ms_per_call_top": {n: round(t, 3) for n, t in top}}


def run_side_c5(args, engine, dsstne_b200, stream, rank, world, local_rank):
    """BASELINE config 5: top-K = 100 over a 1M-item output with the exclusion filter (kCalculateTopK path), batch 4,096.
    N > 1: every rank scores its column shard, the per-rank lists are all-gathered and merged (dsb200_topk_kv) on every rank --
    the scheme of NNNetwork::CalculateTopKGlobal (tests/test_multi_gpu.py checks it against one process)."""
    import torch
    B, N, K = 4096, 1000000, 100
    lo, hi = N * rank // world, N * (rank + 1) // world
    width = hi - lo
    g = torch.Generator(device="cuda").manual_seed(12134 + rank)
    scores = torch.rand(B, width, device="cuda", generator=g)
    # exclusion lists: ~144 columns per row (the user's own history, U/Filters.cpp:49-67), local ids of this shard
    rng = np.random.Generator(np.random.PCG64(7 + rank))
    per = max(1, 144 // world)
    idx = np.sort(rng.integers(0, width, size=(B, per), dtype=np.int64), axis=1).astype(np.uint32).reshape(-1)
    start = (np.arange(B, dtype=np.uint64) * per)
    end = start + np.uint64(per)
    fs = torch.from_numpy(start.view(np.int64)).cuda(); fe = torch.from_numpy(end.view(np.int64)).cuda(); fi = torch.from_numpy(idx.view(np.int32)).cuda()
    ctx = dsstne_b200.Context(local_rank)
    key = torch.empty(B, K, device="cuda"); val = torch.empty(B, K, dtype=torch.int32, device="cuda")
    if world > 1:
        import torch.distributed as dist
        keys = torch.empty(world, B, K, device="cuda"); vals = torch.empty(world, B, K, dtype=torch.int32, device="cuda")
        mk = torch.empty(B, world * K, device="cuda"); mv = torch.empty(B, world * K, dtype=torch.int32, device="cuda")
        ok = torch.empty(B, K, device="cuda"); ov = torch.empty(B, K, dtype=torch.int32, device="cuda")

    def call():
        ctx.topk(scores, K, key, val, filt=(fs, fe, fi))
        if world > 1:
            ctx.topk_offset(val, lo)
            dist.all_gather_into_tensor(keys, key); dist.all_gather_into_tensor(vals, val)
            mk.copy_(keys.permute(1, 0, 2).reshape(B, world * K)); mv.copy_(vals.permute(1, 0, 2).reshape(B, world * K))
            ctx.topk_kv(mk, mv, K, ok, ov)

    for _ in range(2):
        call()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    reps = 3
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(reps):
        call()
    ev1.record()
    torch.cuda.synchronize()
    ms = _max_over_ranks(ev0.elapsed_time(ev1) / reps, world)
    alg = 4 * B * width + 8 * B * K + 4 * len(idx) + 16 * B           # SURVEY 8d, per GPU
    peak, _ = peaks()
    ctx.close()
    return {"workload": "BASELINE config 5: top-K 100 of 4,096 x 1,000,000 scores with exclusion filter (~144 / row)", "metric": "top-K rows/s",
            "value": round(B / (ms * 1e-3), 1), "unit": "rows/s", "ms_per_call": round(ms, 3), "n_gpus": world, "scaling": "strong",
            "roofline": {"bound": "hbm", "achieved": round(alg / (ms * 1e-3) / 1e9, 1), "peak": peak, "unit": "GB/s",
                         "frac": round(alg / (ms * 1e-3) / 1e9 / peak, 4), "algorithmic_bytes_per_gpu": int(alg),
                         "note": "per GPU; N > 1 includes the NCCL all-gather of the per-rank lists and the merge"}}


def run_e2e(args, wl, data, engine, dsstne_b200, stream, world=1):
    """Host-buffer path: per step, the CSR batch goes pinned host -> device through NNDataSet::LoadSparseData
    (the call the reference's JNI binding makes per request) and the loss comes back to the host."""
    import torch
    from dsstne_b200 import datagen
    B = wl["batch"]
    n_batches = data.examples // B
    # carve per-batch CSRs (zero-based starts) out of the synthetic dataset, pinned
    batches = []
    for b in range(n_batches):
        s0 = int(data.start[b * B])
        st = (data.start[b * B:(b + 1) * B] - np.uint64(s0)).astype(np.uint64)
        en = (data.end[b * B:(b + 1) * B] - np.uint64(s0)).astype(np.uint64)
        ix = data.index[s0:int(data.end[(b + 1) * B - 1])]
        batches.append((st, en, np.ascontiguousarray(ix)))
    big = max(batches, key=lambda b: len(b[2]))       # capacity of an NNDataSet is fixed at construction: size it for the largest batch
    first = datagen.HostCsr(big[0], big[1], big[2], wl["items"])
    ds_in = engine.Dataset("gl_input", first.start, first.end, first.index, wl["items"])
    ds_out = engine.Dataset("gl_output", first.start, first.end, first.index, wl["items"])
    net = engine.Network(engine.autoencoder_json(wl["hidden"], smce=SMCE, init=("Gaussian", 0.01, 0.0)), B, [ds_in, ds_out])
    net.set_training_mode(dsstne_b200.SGD)
    net.set_gemm_mode(args.gemm_mode)
    h2d = d2h = 0
    engine.set_option("pinned_mirror", 1 if args.pinned_mirror else 0)

    def step(i):
        nonlocal h2d, d2h
        st, en, ix = batches[i % n_batches]
        ds_in.load_sparse(st, en, ix)
        ds_out.load_sparse(st, en, ix)
        h2d = 2 * (st.nbytes + en.nbytes + ix.nbytes)
        d2h = 8
        return net.train_step(0, HYPER["alpha"], HYPER["lam"], HYPER["lam1"], HYPER["mu"], HYPER["mu1"])

    for i in range(max(args.warmup, 3)):
        step(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(args.steps):
        step(i)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:                                        # every rank loads the same batch and keeps its column shard: slowest rank counts
        import torch.distributed as dist
        t = torch.tensor([dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    net.close()
    return {"value": round(args.steps * B / dt, 1), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
            "ms_per_step": round(dt / args.steps * 1e3, 4),
            "path": "NNDataSet::LoadSparseData (pinned host CSR batch -> device) + NNNetwork train step + loss read-back, every step"}


def cpu_baseline(wl, data, sample_steps=3):
    """OpenMP CPU oracle (oracle/liboracle.so, a restatement -- the reference has no CPU path) on a bounded
    sample of the same workload: `sample_steps` minibatches of the same network and data."""
    from oracle import oracle as orc
    from dsstne_b200 import datagen
    # torch.distributed.run exports OMP_NUM_THREADS=1: the CPU arm uses every core this process may run on
    try:
        orc.set_num_threads(len(os.sched_getaffinity(0)))
    except Exception:
        pass
    B = wl["batch"]
    sizes = [wl["items"]] + wl["hidden"] + [wl["items"]]
    net = orc.Network(sizes, error=orc.ERR_SMCE, mode=orc.SGD, max_batch=B)
    Ws, bs = datagen.make_weights(sizes, scale=0.01)
    for i in range(len(sizes) - 1):
        net.W(i)[:] = Ws[i]
        net.b(i)[:] = bs[i]
    net.s.params = orc.make_params(smce=SMCE)
    sub = datagen.HostCsr(data.start[:B * 2], data.end[:B * 2], data.index[:int(data.end[B * 2 - 1])], wl["items"])
    oc = orc.Csr(sub.start, sub.end, sub.index)
    net.set_input(oc, B)
    net.train_step(oc, oc, 0, B, HYPER["alpha"], HYPER["lam"])          # warm-up (page faults, thread pool)
    t0 = time.perf_counter()
    for i in range(sample_steps):
        net.train_step(oc, oc, (i % 2) * B, B, HYPER["alpha"], HYPER["lam"])
    dt = time.perf_counter() - t0
    net.close()
    return {"value": round(sample_steps * B / dt, 1), "unit": UNIT, "cores": orc.num_threads(), "kind": "port",
            "sample": f"{sample_steps} minibatches of {B} examples of the same workload (after 1 warm-up), OpenMP C oracle, fp32",
            "ms_per_step": round(dt / sample_steps * 1e3, 1)}


def run_reference(args, wl, rank):
    """--impl reference: the CPU arm.  The reference has no CPU implementation of this path and its full build needs
    MPI / NetCDF-C++4 / jsoncpp (absent), so this arm times the CPU oracle port with all host threads."""
    if rank != 0:
        return None
    data = make_data(wl, 2)
    steps = max(1, min(args.steps, 5))
    cb = cpu_baseline(wl, data, sample_steps=steps)
    return {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": 1, "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": {"workload": wl["desc"]},
            "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)            # 0.25 s timed region on c2: several nvidia-smi clock samples fall inside it
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c2", "c4"])
    ap.add_argument("--gemm-mode", type=int, default=2, help="dense GEMMs: 0 exact fp32 (SIMT), 1 tcgen05 TF32, 2 tcgen05 3xTF32 (fp32-grade)")
    ap.add_argument("--gemm-loader", type=int, default=-1, help="operand path of the general tcgen05 kernel (csrc/gemm_tc.cu), -1 = per shape")
    ap.add_argument("--cpu-steps", type=int, default=3)
    ap.add_argument("--side", type=int, default=1, help="1 (default) = after the headline workload also time BASELINE config 4 (1M-item layers) and config 5 (top-K) and add them as \"c4\" / \"c5\" records")
    ap.add_argument("--c4-steps", type=int, default=6)
    ap.add_argument("--fuse-output", type=int, default=1, help="1 (default) = output layer forward GEMM fused with loss + delta (engine option fuse_output_gemm); 0 = two calls")
    ap.add_argument("--pinned-mirror", type=int, default=1, help="e2e path: 1 (default, the engine's default) = LoadSparseData uploads from the page-locked host mirror, one host copy per batch; 0 = two-copy staging path")
    ap.add_argument("--pdl", type=int, default=1, help="1 (default) = programmatic dependent launch of the main-stream kernels (context option pdl)")
    ap.add_argument("--p2p", type=int, default=1, help="N > 1: 1 (default) = exchange steps as one kernel over peer memory each (csrc/comm.cu); 0 = NCCL")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wl = workload(args.workload)
    if args.impl == "reference":
        res = run_reference(args, wl, rank)
    else:
        if world != args.gpus and world == 1 and args.gpus > 1:
            raise SystemExit("bench.py: --gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)")
        res = run_ours(args, wl, rank, world, local_rank)
    if rank == 0 and res is not None:
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
