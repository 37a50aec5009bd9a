"""Generates tests/golden/ref_kernels.npz by running the REFERENCE's own CUDA kernels (compiled unmodified from
/root/reference into oracle/_ref/libdsstne_refkernels.so, see oracle/Makefile) on the seeded inputs of cases.py.
Needs a GPU:   gpurun -- 'python tests/golden/make_ref_golden.py gpurun_out/ref_kernels.npz'
The output is committed; tests/test_oracle_golden.py (CPU) checks the oracle against it."""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import cases  # noqa: E402


def main(out_path):
    import torch
    from oracle import oracle as orc
    ref = orc.ref_kernels()
    assert ref is not None, "oracle/_ref/libdsstne_refkernels.so missing"
    torch.zeros(1, device="cuda")
    assert ref.ref_init() == 0

    keep = []                                              # every device tensor stays alive until the end: the kernels
                                                           # run asynchronously and the caching allocator would reuse the memory
    def dev(a):
        a = np.ascontiguousarray(a)
        if a.dtype == np.uint32:
            a = a.view(np.int32)
        if a.dtype == np.uint64:
            a = a.view(np.int64)
        t = torch.from_numpy(a.copy()).cuda()
        keep.append(t)
        return t

    def p(t):
        return None if t is None else C.c_void_p(t.data_ptr())

    def params(p_den=0.0):
        ref.ref_set_params(C.c_int(0), None, C.c_float(p_den), C.c_float(cases.BOOST[0]), C.c_float(cases.BOOST[1]),
                           C.c_float(cases.SMCE[0]), C.c_float(cases.SMCE[1]), C.c_float(cases.SMCE[2]), C.c_float(cases.SMCE[3]))

    out = {}
    d = cases.dense_inputs(101)
    B, S, N = cases.BATCH, cases.STRIDE, cases.WIDTH
    for name, kw in cases.Z_CASES:
        den = kw.get("denoised", False)
        c = cases.make_csr(7, weighted=kw.get("weighted", False), analog=kw.get("analog", False))
        rnd = d["rnd"][:len(c["index"])] if den else None
        params(cases.DENOISE_P if den else 0.0)
        dz = dev(d["Z0"])
        ref.ref_sparse_z(C.c_uint32(0), C.c_uint32(B), C.c_uint32(S), p(dev(d["W"])), None, p(dev(c["start"])), p(dev(c["end"])),
                         p(dev(c["index"])), p(None if c["weight"] is None else dev(c["weight"])),
                         p(None if c["data"] is None else dev(c["data"])), p(None if rnd is None else dev(rnd)), p(dz), C.c_float(1.0))
        ref.ref_sync()
        out[f"z_{name}"] = dz.cpu().numpy()
    # transposed counts + Boolean gradient
    c = cases.make_csr(7)
    tstart, cap = orc.transposed_capacity(orc.Csr(c["start"], c["end"], c["index"]), N, B)
    params(0.0)
    dstart = dev(tstart)
    dend = dstart.clone()
    didx = torch.zeros(cap, dtype=torch.int32, device="cuda")
    ref.ref_sparse_transpose(C.c_uint32(0), C.c_uint32(B), None, p(dev(c["start"])), p(dev(c["end"])), p(dev(c["index"])), None, None, None,
                             p(dend), p(didx), None)
    dg = torch.zeros((N, S), dtype=torch.float32, device="cuda")
    ref.ref_sparse_wgrad(C.c_float(-1.0 / B), C.c_float(0.0), C.c_uint32(N), C.c_uint32(S), p(dstart), p(dend), p(didx), None, p(dev(d["delta"])), p(dg))
    ref.ref_sync()
    out["t_end"] = dend.cpu().numpy().view(np.uint32)
    tidx = didx.cpu().numpy().view(np.uint32)
    out["t_index_sorted"] = np.concatenate([np.sort(tidx[s:e]) for s, e in zip(tstart, out["t_end"])] + [np.zeros(0, np.uint32)])
    out["wgrad"] = dg.cpu().numpy()
    # activations, losses, deltas
    ref.ref_sparse_loss.restype = C.c_float
    for act in (0, 7):
        da = dev(d["z_out"])
        ref.ref_activation(C.c_int(act), p(da), C.c_uint32(B), C.c_uint32(N), C.c_float(0.0), C.c_float(0.0), C.c_float(0.0))
        ref.ref_sync()
        out[f"act_{act}"] = da.cpu().numpy()
    c = cases.make_csr(11, width=N)
    params(0.0)
    for ef, act, iz in cases.LOSS_CASES:
        unit = orc.activation(act, d["z_out"].copy())
        du = dev(unit)
        loss = ref.ref_sparse_loss(C.c_int(ef), C.c_int(act), C.c_uint32(0), C.c_uint32(B), C.c_uint32(N), p(du), None, p(dev(c["start"])),
                                   p(dev(c["end"])), p(dev(c["index"])), None, C.c_int(int(iz)))
        dd = torch.zeros((B, N), dtype=torch.float32, device="cuda")
        ref.ref_sparse_output_delta(C.c_int(ef), C.c_int(act), C.c_uint32(0), C.c_uint32(B), C.c_uint32(N), p(du), p(dd), None, p(dev(c["start"])),
                                    p(dev(c["end"])), p(dev(c["index"])), None, C.c_int(int(iz)), C.c_float(0.0), C.c_float(0.0), C.c_float(0.0))
        ref.ref_sync()
        out[f"loss_{ef}_{act}_{int(iz)}"] = np.float32(loss)
        out[f"delta_{ef}_{act}_{int(iz)}"] = dd.cpu().numpy()
    # optimizers
    hp = cases.OPT_HP
    for mode in range(7):
        dw, dv, dgv = dev(d["w"]), dev(d["v"]), dev(d["gv"])
        ref.ref_update_weights(C.c_int(mode), C.c_float(hp["alpha"]), C.c_float(hp["lam"]), C.c_float(hp["lam1"]), C.c_float(hp["mu"]),
                               C.c_float(hp["mu1"]), C.c_float(hp["t"]), C.c_uint64(d["w"].size), p(dv), p(dev(d["g"])), p(dgv), p(dw))
        ref.ref_sync()
        out[f"opt_w_{mode}"] = dw.cpu().numpy()
    # top-K
    ok = torch.zeros((B, 64), dtype=torch.float32, device="cuda")
    ov = torch.zeros((B, 64), dtype=torch.int32, device="cuda")
    ref.ref_topk3(p(dev(d["scores"])), p(ok), p(ov), C.c_uint32(B), C.c_uint32(4096), C.c_uint32(64))
    ref.ref_sync()
    out["topk_key"] = ok.cpu().numpy()
    out["topk_val"] = ov.cpu().numpy().view(np.uint32)
    np.savez_compressed(out_path, **out)
    print("wrote", out_path, sum(v.nbytes for v in out.values()), "bytes raw")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "ref_kernels.npz"))
