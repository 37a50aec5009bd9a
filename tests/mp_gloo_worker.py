"""Worker of the CPU model-parallel test (gloo, launched by test_model_parallel_gloo.py under torch.distributed.run).

Runs the sharding scheme of the engine (amazon-dsstne_b200/engine/NNLayer.cpp, which follows E/NNLayer.cpp:1169-1422
forward and :2316-2626 backward, the weight-shard rule of E/NNWeight.cpp:435-457 and the unit ranges of
E/NNLayer.cpp:108-112) with the CPU oracle's kernels in place of the CUDA ones and gloo in place of NCCL, then checks
that the re-assembled result equals the single-process oracle network.  What it covers is the host-side logic of the
N>1 path: who owns which units, which exchange step (reduce-scatter / all-gather) goes where, column-sharded sparse
datasets with local indices, and the all-reduced fixed-point loss."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def shard_csr(orc, h, lo, hi):
    """Column shard of a sparse dataset with indices rebased to the shard (NNDataSet<T>::Shard(Model), E/NNTypes.cpp:2098-2176)."""
    start = np.zeros(len(h.start), np.uint64)
    end = np.zeros(len(h.start), np.uint64)
    rows = []
    pos = 0
    for i, (s, e) in enumerate(zip(h.start, h.end)):
        r = h.index[int(s):int(e)]
        r = r[(r >= lo) & (r < hi)] - np.uint32(lo)
        start[i] = pos
        pos += len(r)
        end[i] = pos
        rows.append(r)
    return orc.Csr(start, end, np.concatenate(rows).astype(np.uint32))


def main():
    import torch
    import torch.distributed as dist
    from dsstne_b200 import datagen
    from oracle import oracle as orc
    from helpers import rel_err, tiny

    sizes = json.loads(os.environ.get("MP_SIZES", "[512, 64, 48, 64, 512]"))
    batch = int(os.environ.get("MP_BATCH", "64"))
    steps = int(os.environ.get("MP_STEPS", "2"))
    alpha, lam = 0.05, 1e-4
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    L = len(sizes) - 1
    rng_ = [orc.shard_range(n, rank, world) for n in sizes]
    loc = [b - a for a, b in rng_]
    outgoing = [orc.weight_outgoing_larger(sizes[i], sizes[i + 1]) for i in range(L)]
    assert not outgoing[0] and outgoing[-1], "test shape must have a column-sharded sparse input and a wider output"

    h = tiny(examples=2 * batch, width=sizes[0])
    Ws, bs = datagen.make_weights(sizes, scale=0.05)
    for i in range(L):
        bs[i][:] = np.random.default_rng(i).standard_normal(bs[i].shape).astype(np.float32) * 0.1
    # local shards: outgoing larger -> full height, output slice; incoming larger -> input slice, full width
    Wl = [(Ws[i][:, rng_[i + 1][0]:rng_[i + 1][1]] if outgoing[i] else Ws[i][rng_[i][0]:rng_[i][1], :]).copy() for i in range(L)]
    bl = [bs[i][rng_[i + 1][0]:rng_[i + 1][1]].copy() for i in range(L)]
    csr_in = shard_csr(orc, h, *rng_[0])
    csr_out = shard_csr(orc, h, *rng_[L])
    params = orc.make_params(smce=(1.0, 0.0, 1.0, 1.0))
    tstart, cap = orc.transposed_capacity(csr_in, loc[0], batch)

    def all_reduce(a):
        t = torch.from_numpy(a)
        dist.all_reduce(t)
        return a

    def reduce_scatter(full, l):                      # NNLayer::Reduce
        all_reduce(full)
        return np.ascontiguousarray(full[:, rng_[l][0]:rng_[l][1]])

    def all_gather(local, l):                         # NNLayer::Gather
        full = np.zeros((batch, sizes[l]), np.float32)
        full[:, rng_[l][0]:rng_[l][1]] = local
        return all_reduce(full)

    losses = []
    for step in range(steps):
        pos = (step % 2) * batch
        X = [None] * (L + 1)
        # ---- forward
        for l in range(1, L + 1):
            i = l - 1
            if outgoing[i]:
                Z = np.zeros((batch, loc[l]), np.float32)
                orc.gemm_fwd(all_gather(X[i], i), Wl[i], Z, beta=0.0)
            else:
                full = np.zeros((batch, sizes[l]), np.float32)
                if i == 0:
                    orc.sparse_z(params, csr_in, pos, batch, Wl[0], full, 1.0)
                else:
                    orc.gemm_fwd(X[i], Wl[i], full, beta=0.0)
                Z = reduce_scatter(full, l)
            Z += bl[i][None, :]
            X[l] = orc.activation(orc.ACT_SIGMOID, np.ascontiguousarray(Z))
        # ---- loss: fixed-point partial sums add exactly across ranks
        part = np.array([orc.sparse_loss(params, csr_out, orc.ERR_SMCE, orc.ACT_SIGMOID, pos, batch, X[L])], np.float64)
        reg = np.array([sum(orc.regularization_error(lam, 0.0, w) for w in Wl)], np.float64)
        losses.append(float(all_reduce(part)[0]) + float(all_reduce(reg)[0]))
        # ---- backward
        D = [None] * (L + 1)
        dW = [np.zeros_like(w) for w in Wl]
        D[L] = orc.sparse_output_delta(params, csr_out, orc.ERR_SMCE, orc.ACT_SIGMOID, pos, batch, X[L], np.zeros_like(X[L]))
        for l in range(L, 0, -1):
            if l < L and outgoing[l]:                 # weights leaving l towards a wider layer
                orc.gemm_dw(all_gather(X[l], l), D[l + 1], dW[l], -1.0 / batch)
                full = np.zeros((batch, sizes[l]), np.float32)
                orc.gemm_dx(D[l + 1], Wl[l], full)
                D[l] = reduce_scatter(full, l)
            if l < L:
                orc.hadamard(orc.ACT_SIGMOID, X[l], D[l])
            i = l - 1
            if not outgoing[i]:                       # weights entering l from a wider layer
                Dfull = all_gather(D[l], l)
                if i == 0:
                    tend, tidx, _ = orc.sparse_transpose(params, csr_in, pos, batch, tstart, cap)
                    orc.sparse_wgrad(params, -1.0 / batch, 0.0, tstart, tend, tidx, None, Dfull, dW[0])
                else:
                    orc.gemm_dw(X[i], Dfull, dW[i], -1.0 / batch)
                    D[i] = np.zeros((batch, loc[i]), np.float32)
                    orc.gemm_dx(Dfull, Wl[i], D[i])
        # ---- update (SGD)
        for i in range(L):
            orc.update_weights(orc.SGD, alpha, lam, 0.0, 0.0, 0.0, 1.0, None, dW[i], None, Wl[i])
            orc.update_biases(orc.SGD, alpha, 0.0, 0.0, 1.0, D[i + 1], None, None, bl[i])

    gathered = [None] * world
    dist.all_gather_object(gathered, (Wl, bl))
    if rank == 0:
        onet = orc.Network(sizes, error=orc.ERR_SMCE, mode=orc.SGD, max_batch=batch)
        for i in range(L):
            onet.W(i)[:] = Ws[i]
            onet.b(i)[:] = bs[i]
        onet.s.params = params
        oc = orc.Csr(h.start, h.end, h.index)
        onet.set_input(oc, batch)
        want = [sum(onet.train_step(oc, oc, (s % 2) * batch, batch, alpha, lam)) for s in range(steps)]
        errs = []
        for i in range(L):
            full = np.zeros_like(Ws[i])
            fb = np.zeros_like(bs[i])
            for r in range(world):
                a0, a1 = orc.shard_range(sizes[i], r, world)
                b0, b1 = orc.shard_range(sizes[i + 1], r, world)
                if outgoing[i]:
                    full[:, b0:b1] = gathered[r][0][i]
                else:
                    full[a0:a1, :] = gathered[r][0][i]
                fb[b0:b1] = gathered[r][1][i]
            errs.append(max(rel_err(full, onet.W(i)), rel_err(fb, onet.b(i))))
        print("RESULT " + json.dumps(dict(losses=losses, want=want, errs=errs, world=world, outgoing=outgoing)))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
