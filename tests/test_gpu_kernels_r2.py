"""GPU parity, second batch (the variants round 1 implemented but never exercised on a GPU):
  * Indexed datasets (NNDataSet::_pbIndex, E/NNTypes.h:527-649 "Indexed" launchers) through the transposed matrix, the loss, the
    output delta and the fused output pass -- sparse Z is covered in test_gpu_kernels.py;
  * `unsigned char` / `char` analog values (scaled by 1/256 and 1/128, E/kernels.cu:923,998) through sparse Z, the transposed
    matrix and the weight gradient;
  * BASELINE config 5 at its stated shape: top-K = 100 of 4,096 x 1,000,000 scores with the exclusion filter, a random subset
    of rows checked against the oracle.
Integer outputs bit-exact, fp32 within 1e-5 (helpers.rel_err)."""
import numpy as np
import pytest

from helpers import ml20m, rel_err, tiny, to_device, to_oracle

pytestmark = pytest.mark.gpu
TOL = 1e-5


def dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.detach().cpu().numpy()


def u32(t):
    return host(t).view(np.uint32)


def indexed(h, examples, seed=5):
    """an example-index table that repeats and permutes the unique examples of h (NNDataSetEnums::Indexed)"""
    rng = np.random.default_rng(seed)
    return rng.integers(0, len(h.start), size=examples).astype(np.uint32)


# ------------------------------------------------------------------ Indexed: transposed matrix
@pytest.mark.parametrize("weighted", [False, True])
def test_indexed_sparse_transpose_bit_exact(ctx, orc, dsb, weighted):
    import torch
    h = tiny(examples=200, width=2048, weighted=weighted)
    ex = indexed(h, 256)
    batch = 256
    oc = to_oracle(orc, h, ex_index=ex)
    tstart, cap = orc.transposed_capacity(oc, h.width, batch)
    params = orc.make_params()
    r_end, r_idx, r_data = orc.sparse_transpose(params, oc, 0, batch, tstart, cap, False)
    ctx.set_params()
    d_start = torch.from_numpy(tstart.view(np.int32).copy()).cuda()
    d_end = torch.zeros(h.width, dtype=torch.int32, device="cuda")
    d_idx = torch.zeros(max(cap, 1), dtype=torch.int32, device="cuda")
    d_data = torch.zeros(max(cap, 1), dtype=torch.float32, device="cuda") if r_data is not None else None
    ctx.sparse_transpose(to_device(dsb, h, ex_index=ex), 0, batch, h.width, d_start, d_end, d_idx, d_data, False)
    ctx.sync()
    np.testing.assert_array_equal(u32(d_end), r_end)
    g_idx = u32(d_idx)
    for c in np.nonzero(r_end > tstart)[0]:
        s, e = tstart[c], r_end[c]
        np.testing.assert_array_equal(g_idx[s:e], r_idx[s:e])
        if r_data is not None:
            np.testing.assert_array_equal(host(d_data)[s:e], r_data[s:e])


# ------------------------------------------------------------------ Indexed: loss, delta, fused pass
@pytest.mark.parametrize("ef", [3, 2, 1], ids=["smce", "ce", "l2"])
@pytest.mark.parametrize("weighted", [False, True])
def test_indexed_loss_delta_and_fused_pass(ctx, orc, dsb, ef, weighted):
    import torch
    batch, stride = 128, 2050
    h = tiny(examples=100, width=stride, weighted=weighted)
    ex = indexed(h, batch)
    rng = np.random.default_rng(3)
    z = (rng.standard_normal((batch, stride)) * 2.0 - 1.0).astype(np.float32)
    unit = orc.activation(orc.ACT_SIGMOID, z.copy())
    smce = (0.8, 0.05, 1.5, 0.75)
    params = orc.make_params(smce=smce)
    oc = to_oracle(orc, h, ex_index=ex)
    ref_loss = orc.sparse_loss(params, oc, ef, orc.ACT_SIGMOID, 0, batch, unit)
    ref_delta = orc.sparse_output_delta(params, oc, ef, orc.ACT_SIGMOID, 0, batch, unit, np.zeros_like(unit))
    ctx.set_params(smce=smce)
    try:
        dd = to_device(dsb, h, ex_index=ex)
        d_unit = dev(unit)
        got_loss = ctx.sparse_loss(dd, ef, dsb.ACT_SIGMOID, 0, batch, d_unit)
        d_delta = dev(np.full_like(unit, 7.0))
        ctx.sparse_output_delta(dd, ef, dsb.ACT_SIGMOID, 0, batch, d_unit, d_delta)
        f_delta = torch.empty_like(d_delta)
        acc = torch.zeros(1, dtype=torch.int64, device="cuda")
        ctx.output_pass(dd, ef, dsb.ACT_SIGMOID, 0, batch, dev(z), None, f_delta, acc)
        ctx.sync()
    finally:
        ctx.set_params()
    assert abs(got_loss - ref_loss) <= TOL * max(abs(ref_loss), 1.0)
    assert rel_err(host(d_delta), ref_delta) < TOL
    assert rel_err(host(f_delta), ref_delta) < TOL
    assert abs(float(acc.item()) / float(1 << 30) - ref_loss) <= TOL * max(abs(ref_loss), 1.0)


def test_indexed_fused_output_gemm(ctx, orc, dsb):
    """the tcgen05 fused forward (csrc/gemm_stream.cu) builds its target bitmap through the example index as well"""
    import torch
    B, k, n = 128, 64, 3000
    h = tiny(examples=90, width=n, weighted=True)
    ex = indexed(h, B)
    g = torch.Generator(device="cuda").manual_seed(9)
    A = torch.rand(B, k, device="cuda", generator=g)
    W = torch.randn(k, n, device="cuda", generator=g) * 0.1
    bias = torch.randn(n, device="cuda", generator=g) * 0.5 - 1.0
    z = (A.double() @ W.double() + bias.double()).float().cpu().numpy()
    unit = orc.activation(orc.ACT_SIGMOID, z.copy())
    params = orc.make_params(smce=(1.0, 0.0, 1.0, 1.0))
    oc = to_oracle(orc, h, ex_index=ex)
    ref_loss = orc.sparse_loss(params, oc, 3, orc.ACT_SIGMOID, 0, B, unit)
    ref_delta = orc.sparse_output_delta(params, oc, 3, orc.ACT_SIGMOID, 0, B, unit, np.zeros_like(unit))
    ctx.set_params(smce=(1.0, 0.0, 1.0, 1.0))
    ctx.set_option("gemm_mode", 2)
    try:
        delta = torch.empty(B, n, device="cuda")
        acc = torch.zeros(1, dtype=torch.int64, device="cuda")
        ctx.gemm_fwd_output_pass(to_device(dsb, h, ex_index=ex), 3, dsb.ACT_SIGMOID, 0, A, W, bias, None, delta, acc)
        ctx.sync()
    finally:
        ctx.set_option("gemm_mode", 0)
        ctx.set_params()
    assert rel_err(host(delta), ref_delta) < 3e-5                         # tensor-core bound (tests/test_gpu_gemm.py)
    assert abs(float(acc.item()) / float(1 << 30) - ref_loss) <= 3e-5 * abs(ref_loss)


# ------------------------------------------------------------------ unsigned char / char analog values
@pytest.mark.parametrize("dtype", [np.uint8, np.int8], ids=["uchar", "char"])
def test_byte_valued_analog_data(ctx, orc, dsb, dtype):
    """sparse Z, transposed matrix and weight gradient with 8-bit analog values: uchar * 1/256, char * 1/128"""
    import torch
    batch, stride = 256, 128
    h = tiny(examples=batch, width=2048)
    rng = np.random.default_rng(11)
    h.data = (rng.integers(1, 255, size=h.nnz).astype(np.uint8) if dtype == np.uint8
              else rng.integers(-127, 127, size=h.nnz).astype(np.int8))
    W = (rng.standard_normal((h.width, stride)) * 0.05).astype(np.float32)
    Z0 = rng.standard_normal((batch, stride)).astype(np.float32)
    params = orc.make_params()
    oc = to_oracle(orc, h)
    ref_z = orc.sparse_z(params, oc, 0, batch, W, Z0.copy(), 1.0, False)
    ctx.set_params()
    dd = to_device(dsb, h)
    dZ = dev(Z0)
    ctx.sparse_z(dd, 0, batch, dev(W), dZ, 1.0, False)
    ctx.sync()
    assert rel_err(host(dZ), ref_z) < TOL
    # transposed matrix carries the scaled values
    tstart, cap = orc.transposed_capacity(oc, h.width, batch)
    r_end, r_idx, r_data = orc.sparse_transpose(params, oc, 0, batch, tstart, cap, False)
    d_start = torch.from_numpy(tstart.view(np.int32).copy()).cuda()
    d_end = torch.zeros(h.width, dtype=torch.int32, device="cuda")
    d_idx = torch.zeros(max(cap, 1), dtype=torch.int32, device="cuda")
    d_data = torch.zeros(max(cap, 1), dtype=torch.float32, device="cuda")
    ctx.sparse_transpose(dd, 0, batch, h.width, d_start, d_end, d_idx, d_data, False)
    ctx.sync()
    np.testing.assert_array_equal(u32(d_end), r_end)
    for c in np.nonzero(r_end > tstart)[0]:
        s, e = tstart[c], r_end[c]
        np.testing.assert_array_equal(u32(d_idx)[s:e], r_idx[s:e])
        np.testing.assert_array_equal(host(d_data)[s:e], r_data[s:e])
    # and the gradient built from it (fixed point: bit exact)
    delta = (rng.standard_normal((batch, stride)) * 0.1).astype(np.float32)
    ref_dw = orc.sparse_wgrad(params, -1.0 / batch, 0.0, tstart, r_end, r_idx, r_data, delta, np.zeros((h.width, stride), dtype=np.float32))
    dW = torch.zeros(h.width, stride, device="cuda")
    ctx.sparse_wgrad(-1.0 / batch, 0.0, d_start, d_end, d_idx, d_data, dev(delta), dW)
    ctx.sync()
    np.testing.assert_array_equal(host(dW), ref_dw)


# ------------------------------------------------------------------ BASELINE config 5 at its stated shape
def test_topk_config5_shape_row_subset(ctx, orc, dsb):
    """K = 100 of 4,096 x 1,000,000 scores (16.4 GB) with ~144 excluded columns per row; 24 rows checked against the oracle
    (descending score, ties by ascending column)."""
    import torch
    B, N, K = 4096, 1000000, 100
    g = torch.Generator(device="cuda").manual_seed(5)
    scores = torch.rand(B, N, device="cuda", generator=g)
    h = ml20m(examples=B, width=N)
    dcsr = to_device(dsb, h)
    ok = torch.empty((B, K), dtype=torch.float32, device="cuda")
    ov = torch.empty((B, K), dtype=torch.int32, device="cuda")
    ctx.topk(scores, K, ok, ov, filt=(dcsr.start, dcsr.end, dcsr.index))
    ctx.sync()
    rows = np.random.default_rng(1).choice(B, size=24, replace=False)
    for r in rows:
        s = scores[r].cpu().numpy()[None, :]
        fs = np.array([0], dtype=np.uint64)
        fe = np.array([h.end[r] - h.start[r]], dtype=np.uint64)
        fi = h.index[int(h.start[r]):int(h.end[r])]
        ref_k, ref_v = orc.topk(np.ascontiguousarray(s), K, filt=(fs, fe, fi))
        np.testing.assert_array_equal(host(ok[r]), ref_k[0])
        np.testing.assert_array_equal(u32(ov[r]), ref_v[0])
    del scores
    torch.cuda.empty_cache()


# ------------------------------------------------------------------ small dense layer: gradient + optimizer + bias in one launch
@pytest.mark.parametrize("mode", range(7), ids=["SGD", "Momentum", "AdaGrad", "Nesterov", "RMSProp", "AdaDelta", "Adam"])
@pytest.mark.parametrize("B,k,n", [(1024, 128, 128), (256, 100, 37), (1000, 64, 200)])
def test_dense_update_equals_the_three_calls(ctx, dsb, mode, B, k, n):
    """dsb200_dense_update against dsb200_gemm_dw (fp32) + dsb200_update_weights + dsb200_update_biases on the same inputs"""
    import torch
    g = torch.Generator(device="cuda").manual_seed(100 * mode + B)
    X = torch.rand(B, k, device="cuda", generator=g)
    D = torch.randn(B, n, device="cuda", generator=g) * 0.1
    W0 = torch.randn(k, n, device="cuda", generator=g) * 0.05
    b0 = torch.randn(n, device="cuda", generator=g) * 0.1
    V0 = torch.rand(k, n, device="cuda", generator=g) * 0.01
    GV0 = torch.rand(k, n, device="cuda", generator=g) * 0.01 + 1e-3
    bV0 = torch.rand(n, device="cuda", generator=g) * 0.01
    bGV0 = torch.rand(n, device="cuda", generator=g) * 0.01 + 1e-3
    hp = dict(alpha=0.05, lam=1e-3, lam1=1e-4, mu=0.9, mu1=0.999, t=3.0)
    galpha = -1.0 / B
    # reference: the three calls
    Wr, br, Vr, GVr, bVr, bGVr = W0.clone(), b0.clone(), V0.clone(), GV0.clone(), bV0.clone(), bGV0.clone()
    G = torch.zeros(k, n, device="cuda")
    ctx.set_option("gemm_mode", 0)
    ctx.gemm_dw(X, D, G, galpha)
    ctx.update_weights(mode, hp["alpha"], hp["lam"], hp["lam1"], hp["mu"], hp["mu1"], hp["t"], Vr, G, GVr, Wr)
    ctx.update_biases(mode, hp["alpha"], hp["mu"], hp["mu1"], hp["t"], D, bVr, bGVr, br)
    # one launch
    W, b, V, GV, bV, bGV = W0.clone(), b0.clone(), V0.clone(), GV0.clone(), bV0.clone(), bGV0.clone()
    ctx.dense_update(mode, galpha, X, D, hp["alpha"], hp["lam"], hp["lam1"], hp["mu"], hp["mu1"], hp["t"], V, GV, W, bV, bGV, b)
    ctx.sync()
    tol = 2e-5 if mode in (dsb.ADAGRAD, dsb.RMSPROP, dsb.ADADELTA, dsb.ADAM) else TOL     # rsqrt of a sum of squares amplifies the GEMM rounding
    assert rel_err(host(W), host(Wr)) < tol
    assert rel_err(host(b), host(br)) < tol
    if mode != dsb.SGD:
        assert rel_err(host(V), host(Vr)) < tol and rel_err(host(bV), host(bVr)) < tol
    if mode in (dsb.ADADELTA, dsb.ADAM):
        assert rel_err(host(GV), host(GVr)) < tol and rel_err(host(bGV), host(bGVr)) < tol
