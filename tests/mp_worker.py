"""Worker of the multi-GPU tests: one rank per GPU under torch.distributed.run.  Builds the same sparse
autoencoder on every rank (model parallel: each rank holds its unit slice of every layer / weight), runs a few
training steps through the C++ engine with NCCL exchange steps, and rank 0 compares the re-assembled weights and
the losses with the single-process CPU oracle.  Result -> JSON on stdout (rank 0)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    import dsstne_b200
    from dsstne_b200 import datagen, engine
    from oracle import oracle as orc
    from helpers import rel_err, tiny

    sizes = json.loads(os.environ.get("MP_SIZES", "[2048, 128, 64, 128, 2048]"))
    batch = int(os.environ.get("MP_BATCH", "256"))
    mode = int(os.environ.get("MP_MODE", "0"))
    steps = int(os.environ.get("MP_STEPS", "3"))
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        buf = torch.tensor(list(engine.unique_id()), dtype=torch.uint8, device="cuda")
    dist.broadcast(buf, 0)
    engine.startup(rank, world, local, bytes(buf.cpu().tolist()), seed=12134)
    engine.use_torch_stream(local)
    # exchange steps: one kernel over peer memory each (csrc/comm.cu, the default) or NCCL (MP_P2P=0)
    engine.set_option("p2p_exchange", int(os.environ.get("MP_P2P", "1")))
    gemm_mode = int(os.environ.get("MP_GEMM", "0"))   # 2 = tcgen05 3xTF32: the fused output-layer forward + streamed dW / dX on the shards

    if os.environ.get("MP_DATA") == "ml20m":
        from helpers import ml20m
        h = ml20m(examples=2 * batch, width=sizes[0])
    else:
        h = tiny(examples=2 * batch, width=sizes[0])
    ds_in = engine.Dataset.from_host_csr("gl_input", h)
    ds_out = engine.Dataset.from_host_csr("gl_output", h)
    hidden = sizes[1:-1]
    net = engine.Network(engine.autoencoder_json(hidden, sparseness=(0.5, 2.0)), batch, [ds_in, ds_out])
    net.set_training_mode(mode)
    net.set_gemm_mode(gemm_mode)
    Ws, bs = datagen.make_weights(sizes, scale=0.05)
    names = ["Input"] + [f"Hidden{i + 1}" for i in range(len(hidden))] + ["Output"]
    for i in range(len(sizes) - 1):
        bs[i][:] = np.random.default_rng(i).standard_normal(bs[i].shape).astype(np.float32) * 0.1
        net.set_weights(names[i], names[i + 1], Ws[i], bs[i])
    hp = dict(alpha=0.025, lam=1e-4, lam1=0.0, mu=0.5, mu1=0.999)
    losses = [net.train_step((s % 2) * batch, hp["alpha"], hp["lam"], hp["lam1"], hp["mu"], hp["mu1"]) for s in range(steps)]

    # streaming path while model parallel: every rank re-loads the whole dataset with its two halves exchanged
    # (NNDataSet::LoadSparseData re-slices the rank's column shard) and runs one more step at position 0
    lens = (h.end - h.start).astype(np.uint64)
    order = np.concatenate([np.arange(batch, 2 * batch), np.arange(0, batch)])
    s_end = np.cumsum(lens[order]).astype(np.uint64)
    s_start = np.concatenate([[0], s_end[:-1]]).astype(np.uint64)
    s_index = np.concatenate([h.index[int(h.start[r]):int(h.end[r])] for r in order]).astype(np.uint32)
    ds_in.load_sparse(s_start, s_end, s_index)
    ds_out.load_sparse(s_start, s_end, s_index)
    stream_loss = net.train_step(0, hp["alpha"], hp["lam"], hp["lam1"], hp["mu"], hp["mu1"])

    # model-parallel top-K (MP_TOPK = K): the engine's own output scores, re-assembled from the shards, give the expected global
    # top-K through the oracle; NNNetwork::CalculateTopKGlobal must return exactly that on EVERY rank (exclusion filter = input set)
    topk, topk_ok = int(os.environ.get("MP_TOPK", "0")), None
    if topk:
        net.set_position(0)
        net.predict_batch()
        # the unit buffer is sized for the LARGEST shard (E/NNLayer.cpp:113, as the reference allocates it): the rank's rows are the
        # first batch x localStride floats
        _, local_stride, _, _ = net.layer_info("Output")
        units = np.asarray(net.get_units("Output"), dtype=np.float32)[:batch * local_stride].reshape(batch, local_stride)
        got_k, got_v = net.topk_global("Output", topk, batch, filt=ds_in)
        parts = [None] * world
        dist.all_gather_object(parts, (units, got_k, got_v))
        if rank == 0:
            full = np.ascontiguousarray(np.concatenate([p[0] for p in parts], axis=1))
            sl = slice(0, batch)
            want_k, want_v = orc.topk(full, topk, filt=(s_start[sl], s_end[sl], s_index))
            topk_ok = all(np.array_equal(p[1], want_k) and np.array_equal(p[2], want_v) for p in parts)
            if not topk_ok:                                      # which rank, which row, what differs (shown by the test on failure)
                for r, p in enumerate(parts):
                    bad = np.nonzero((p[1] != want_k).any(axis=1) | (p[2] != want_v).any(axis=1))[0]
                    if bad.size:
                        b = int(bad[0])
                        print(f"MP_TOPK_DIFF rank {r}: {bad.size} rows differ; row {b}: got keys {p[1][b][:6]} ids {p[2][b][:6]} want keys {want_k[b][:6]} ids {want_v[b][:6]}",
                              file=sys.stderr, flush=True)

    # re-assemble the sharded weights on rank 0
    shards = []
    for i in range(len(sizes) - 1):
        W, b = net.get_weights(names[i], names[i + 1])
        shards.append((W, b))
    gathered = [None] * world
    dist.all_gather_object(gathered, shards)
    out = None
    if rank == 0:
        onet = orc.Network(sizes, error=orc.ERR_SMCE, mode=mode, max_batch=batch)
        for i in range(len(sizes) - 1):
            onet.W(i)[:] = Ws[i]
            onet.b(i)[:] = bs[i]
        onet.s.params = orc.make_params(smce=(1.0, 0.0, 1.0, 1.0))
        onet.s.sparsenessPenalty_p, onet.s.sparsenessPenalty_beta = 0.5, 2.0
        for l in range(1, len(sizes) - 1):
            onet.s.sparsePenalty[l] = 1
        oc = orc.Csr(h.start, h.end, h.index)
        onet.set_input(oc, batch)
        want_losses = [onet.train_step(oc, oc, (s % 2) * batch, batch, hp["alpha"], hp["lam"], hp["lam1"], hp["mu"], hp["mu1"])[0]
                       for s in range(steps)]
        want_stream = onet.train_step(oc, oc, batch, batch, hp["alpha"], hp["lam"], hp["lam1"], hp["mu"], hp["mu1"])[0]   # the exchanged first half
        errs = {}
        for i in range(len(sizes) - 1):
            n_in, n_out = sizes[i], sizes[i + 1]
            outgoing = orc.weight_outgoing_larger(n_in, n_out)
            full = np.zeros((n_in, n_out), dtype=np.float32)
            fullb = np.zeros(n_out, dtype=np.float32)
            for r in range(world):
                W, b = gathered[r][i]
                o0, o1 = orc.shard_range(n_out, r, world)
                fullb[o0:o1] = b
                if outgoing:
                    full[:, o0:o1] = W.reshape(n_in, o1 - o0)
                else:
                    i0, i1 = orc.shard_range(n_in, r, world)
                    full[i0:i1, :] = W.reshape(i1 - i0, n_out)
            errs[f"W{i}"] = rel_err(full, onet.W(i))
            errs[f"b{i}"] = rel_err(fullb, onet.b(i))
        out = {"world": world, "losses": losses, "want_losses": want_losses, "errs": errs,
               "loss_err": max(abs(a - b) / abs(b) for a, b in zip(losses, want_losses)),
               "stream_loss_err": abs(stream_loss - want_stream) / abs(want_stream), "topk_ok": topk_ok}
    net.close()
    dist.barrier()
    dist.destroy_process_group()
    engine.shutdown()
    if rank == 0:
        print("MP_RESULT " + json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
