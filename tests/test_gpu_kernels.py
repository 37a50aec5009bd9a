"""GPU parity: every kernel family of the C ABI (include/dsstne_b200.h) against the CPU oracle
on the same seeded inputs.  Integer outputs must be bit-exact; fp32 outputs within 1e-5 relative
(the tolerance BASELINE.json:north_star states), measured against max(|ref|, 1e-3*max|ref|).
"""
import os

import numpy as np
import pytest

from helpers import canon_columns, ml20m, rel_err, tiny, to_device, to_oracle, with_long_rows

pytestmark = pytest.mark.gpu
TOL = 1e-5


def dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.detach().cpu().numpy()


def u32(t):
    return host(t).view(np.uint32)


# ------------------------------------------------------------------ a14
def test_clear_unit_add_bias(ctx, orc):
    import torch
    rng = np.random.default_rng(1)
    bias = rng.standard_normal(130).astype(np.float32)
    unit = rng.standard_normal((37, 130)).astype(np.float32)
    d_unit = dev(unit)
    ctx.add_bias(d_unit, dev(bias))
    np.testing.assert_array_equal(host(d_unit), unit + bias[None, :])
    ctx.clear_unit(d_unit, dev(bias))
    np.testing.assert_array_equal(host(d_unit), np.tile(bias, (37, 1)))
    torch.cuda.synchronize()


# ------------------------------------------------------------------ a1-a3
@pytest.fixture(params=["warp", "staged"])
def zkernel(request, ctx):
    """both sparse-Z kernels: the warp-autonomous one (default for stride % 4 == 0) and the TMA-staged CTA kernel"""
    ctx.set_option("z_staged_kernel", int(request.param == "staged"))
    yield request.param
    ctx.set_option("z_staged_kernel", 0)


def _run_sparse_z(ctx, orc, dsb, h, stride, batch, position=0, beta=1.0, denoised=False, shuffle=None,
                  ex_index=None, p=0.0, no_tma=False):
    import torch
    rng = np.random.default_rng(12345)
    W = (rng.standard_normal((h.width, stride)) * 0.05).astype(np.float32)
    Z0 = rng.standard_normal((batch, stride)).astype(np.float32)
    rnd = rng.random(h.nnz).astype(np.float32) if denoised else None
    params = orc.make_params(shuffle=shuffle, denoising_p=p)
    ref = orc.sparse_z(params, to_oracle(orc, h, random=rnd, ex_index=ex_index), position, batch, W, Z0.copy(), beta, denoised)
    d_shuffle = None if shuffle is None else torch.from_numpy(shuffle.view(np.int32).copy()).cuda()
    ctx.set_params(shuffle=d_shuffle, denoising_p=p)
    ctx.set_option("no_tma", int(no_tma))
    dZ = dev(Z0)
    ctx.sparse_z(to_device(dsb, h, random=rnd, ex_index=ex_index), position, batch, dev(W), dZ, beta, denoised)
    ctx.sync()
    ctx.set_option("no_tma", 0)
    ctx.set_params()
    return host(dZ), ref


@pytest.mark.parametrize("stride", [128, 1024, 64, 96, 100, 2048])
def test_sparse_z_boolean_strides(zkernel, ctx, orc, dsb, stride):
    h = tiny(examples=256)
    got, ref = _run_sparse_z(ctx, orc, dsb, h, stride, 256)
    assert rel_err(got, ref) < TOL


@pytest.mark.parametrize("beta", [0.0, 1.0, 0.5])
def test_sparse_z_beta_and_empty_rows(zkernel, ctx, orc, dsb, beta):
    h = tiny(examples=300, empty_rows=17)
    got, ref = _run_sparse_z(ctx, orc, dsb, h, 128, 256, position=44, beta=beta)
    assert rel_err(got, ref) < TOL


def test_sparse_z_ml20m_shape_long_rows(zkernel, ctx, orc, dsb):
    import scipy.sparse as sp
    h = with_long_rows(ml20m(examples=1024), [9254, 4609, 4608, 513, 512, 1])
    got, ref = _run_sparse_z(ctx, orc, dsb, h, 128, 1024)
    # a 9,254-term serial fp32 sum (the oracle, like the reference kernel) carries more rounding
    # error than 1e-5; judge both against the float64 result instead
    rng = np.random.default_rng(12345)                       # same draws as _run_sparse_z
    W = (rng.standard_normal((h.width, 128)) * 0.05).astype(np.float32)
    Z0 = rng.standard_normal((1024, 128)).astype(np.float32)
    indptr = np.concatenate([h.start, h.end[-1:]]).astype(np.int64)
    A = sp.csr_matrix((np.ones(h.nnz), h.index.astype(np.int64), indptr), shape=(1024, h.width))
    exact = Z0.astype(np.float64) + A @ W.astype(np.float64)
    assert rel_err(got, exact) < TOL
    assert rel_err(ref, exact) < 1e-4
    got2, _ = _run_sparse_z(ctx, orc, dsb, h, 128, 1024)
    np.testing.assert_array_equal(got, got2)          # split-row combine is deterministic


def test_sparse_z_plain_load_path_matches_tma(zkernel, ctx, orc, dsb):
    h = with_long_rows(ml20m(examples=512), [3000, 700])
    a, ref = _run_sparse_z(ctx, orc, dsb, h, 128, 512, no_tma=False)
    b, _ = _run_sparse_z(ctx, orc, dsb, h, 128, 512, no_tma=True)
    np.testing.assert_array_equal(a, b)
    assert rel_err(a, ref) < TOL


def test_sparse_z_analog_weighted(zkernel, ctx, orc, dsb):
    h = tiny(examples=256, analog=True, weighted=True)
    got, ref = _run_sparse_z(ctx, orc, dsb, h, 128, 256)
    assert rel_err(got, ref) < TOL


@pytest.mark.parametrize("analog", [False, True])
def test_sparse_z_denoised(zkernel, ctx, orc, dsb, analog):
    h = tiny(examples=256, analog=analog, weighted=True)
    got, ref = _run_sparse_z(ctx, orc, dsb, h, 128, 256, denoised=True, p=0.2)
    assert rel_err(got, ref) < TOL


def test_sparse_z_shuffled_indexed(zkernel, ctx, orc, dsb):
    h = tiny(examples=200)
    rng = np.random.default_rng(3)
    ex_index = rng.integers(0, 200, size=400).astype(np.uint32)      # Indexed: 400 examples over 200 rows
    shuffle = rng.permutation(400).astype(np.uint32)
    got, ref = _run_sparse_z(ctx, orc, dsb, h, 128, 256, position=100, shuffle=shuffle, ex_index=ex_index)
    assert rel_err(got, ref) < TOL


def test_sparse_z_bias_act_fused(zkernel, ctx, orc, dsb):
    h = tiny(examples=256, empty_rows=9)
    rng = np.random.default_rng(5)
    W = (rng.standard_normal((h.width, 128)) * 0.05).astype(np.float32)
    bias = rng.standard_normal(128).astype(np.float32)
    ref = np.zeros((256, 128), dtype=np.float32)
    orc.clear_unit(ref, bias)
    orc.sparse_z(orc.make_params(), to_oracle(orc, h), 0, 256, W, ref, 1.0)
    orc.activation(orc.ACT_SIGMOID, ref)
    out = dev(np.zeros((256, 128), dtype=np.float32))
    ctx.sparse_z_bias_act(to_device(dsb, h), 0, 256, dev(W), dev(bias), dsb.ACT_SIGMOID, out)
    ctx.sync()
    assert rel_err(host(out), ref) < TOL


# ------------------------------------------------------------------ a4-a6
def _transpose_case(ctx, orc, dsb, h, batch, position=0, denoised=False, p=0.0, sort=True):
    import torch
    rng = np.random.default_rng(99)
    rnd = rng.random(h.nnz).astype(np.float32) if denoised else None
    oc = to_oracle(orc, h, random=rnd)
    tstart, cap = orc.transposed_capacity(oc, h.width, batch)
    params = orc.make_params(denoising_p=p)
    r_end, r_idx, r_data = orc.sparse_transpose(params, oc, position, batch, tstart, cap, denoised)
    ctx.set_params(denoising_p=p)
    ctx.set_option("transpose_sort", int(sort))
    d_start = torch.from_numpy(tstart.view(np.int32).copy()).cuda()
    d_end = torch.zeros(h.width, dtype=torch.int32, device="cuda")
    d_idx = torch.zeros(max(cap, 1), dtype=torch.int32, device="cuda")
    d_data = torch.zeros(max(cap, 1), dtype=torch.float32, device="cuda") if r_data is not None else None
    ctx.sparse_transpose(to_device(dsb, h, random=rnd), position, batch, h.width, d_start, d_end, d_idx, d_data, denoised)
    ctx.sync()
    ctx.set_option("transpose_sort", 1)
    return (tstart, r_end, r_idx, r_data), (d_start, d_end, d_idx, d_data), params


@pytest.mark.parametrize("case", ["tiny", "ml20m", "weighted", "analog", "denoised", "weighted_denoised"])
def test_sparse_transpose_bit_exact(ctx, orc, dsb, case):
    if case == "tiny":
        h, batch = tiny(256), 256
    elif case == "ml20m":
        h, batch = with_long_rows(ml20m(1024), [5000]), 1024
    elif case == "weighted":
        h, batch = tiny(256, weighted=True), 256
    elif case == "analog":
        h, batch = tiny(256, analog=True, weighted=True), 256
    else:
        h, batch = tiny(256, weighted=(case == "weighted_denoised")), 256
    den = case.endswith("denoised")
    (tstart, r_end, r_idx, r_data), (d_start, d_end, d_idx, d_data), _ = _transpose_case(
        ctx, orc, dsb, h, batch, denoised=den, p=0.25 if den else 0.0)
    g_end, g_idx = u32(d_end), u32(d_idx)
    np.testing.assert_array_equal(g_end, r_end)                        # counts: bit exact
    # the kernel emits the canonical (ascending row) order itself: compare raw arrays
    for c in np.nonzero(r_end > tstart)[0]:
        s, e = tstart[c], r_end[c]
        np.testing.assert_array_equal(g_idx[s:e], r_idx[s:e])
        if r_data is not None:
            np.testing.assert_array_equal(host(d_data)[s:e], r_data[s:e])


def test_sparse_transpose_unsorted_is_same_set(ctx, orc, dsb):
    h = ml20m(512)
    (tstart, r_end, r_idx, _), (_, d_end, d_idx, _), _ = _transpose_case(ctx, orc, dsb, h, 512, sort=False)
    np.testing.assert_array_equal(u32(d_end), r_end)
    got = canon_columns(tstart, r_end, u32(d_idx))
    ref = canon_columns(tstart, r_end, r_idx)
    for a, b in zip(got, ref):
        np.testing.assert_array_equal(a, b)


@pytest.fixture(params=["unified", "two_kernel", "tile"])
def gkernel(request, ctx):
    """the three sparse-gradient schemes: the unified kernel (heavy-column items + light columns in one grid; default when
    n % 128 == 0), round 1's warp-per-light-column + CTA-per-heavy-column pair, and the tile kernel"""
    ctx.set_option("wgrad_tile_kernel", int(request.param == "tile"))
    ctx.set_option("wgrad_two_kernel", int(request.param == "two_kernel"))
    yield request.param
    ctx.set_option("wgrad_tile_kernel", 0)
    ctx.set_option("wgrad_two_kernel", 0)


@pytest.mark.parametrize("case", ["tiny128", "ml20m128", "tiny1024", "analog", "beta", "stride100"])
def test_sparse_wgrad_bit_exact(gkernel, ctx, orc, dsb, case):
    import torch
    n = {"tiny1024": 1024, "stride100": 100}.get(case, 128)
    if case == "ml20m128":
        h, batch = ml20m(1024), 1024
    elif case == "analog":
        h, batch = tiny(256, analog=True, weighted=True), 256
    else:
        h, batch = tiny(256), 256
    (tstart, r_end, r_idx, r_data), (d_start, d_end, d_idx, d_data), params = _transpose_case(ctx, orc, dsb, h, batch)
    rng = np.random.default_rng(4)
    delta = (rng.standard_normal((batch, n)) * 0.1).astype(np.float32)
    beta = 0.5 if case == "beta" else 0.0
    dW0 = rng.standard_normal((h.width, n)).astype(np.float32)
    alpha = -1.0 / batch
    ref = orc.sparse_wgrad(params, alpha, beta, tstart, r_end, r_idx, r_data, delta, dW0.copy())
    d_dW = dev(dW0)
    ctx.sparse_wgrad(alpha, beta, d_start, d_end, d_idx, d_data, dev(delta), d_dW)
    ctx.sync()
    got = host(d_dW)
    # fixed-point sums are order independent: identical integers -> identical floats
    np.testing.assert_array_equal(got, ref)


@pytest.mark.parametrize("mode", range(7))
def test_sparse_wgrad_update_fused_equals_unfused(gkernel, ctx, orc, dsb, mode):
    h, batch, n = tiny(256), 256, 128
    (tstart, r_end, r_idx, r_data), (d_start, d_end, d_idx, d_data), params = _transpose_case(ctx, orc, dsb, h, batch)
    rng = np.random.default_rng(8)
    delta = (rng.standard_normal((batch, n)) * 0.1).astype(np.float32)
    w0 = (rng.standard_normal((h.width, n)) * 0.05).astype(np.float32)
    v0 = (rng.random((h.width, n)) * 0.01).astype(np.float32)
    gv0 = (rng.random((h.width, n)) * 0.01).astype(np.float32)
    g = orc.sparse_wgrad(params, -1.0 / batch, 0.0, tstart, r_end, r_idx, r_data, delta, np.zeros_like(w0))
    w_ref, v_ref, gv_ref = w0.copy(), v0.copy(), gv0.copy()
    hp = dict(alpha=0.025, lam=1e-4, lam1=1e-5, mu=0.9, mu1=0.999, t=3.0)
    orc.update_weights(mode, hp["alpha"], hp["lam"], hp["lam1"], hp["mu"], hp["mu1"], hp["t"], v_ref, g, gv_ref, w_ref)
    d_w, d_v, d_gv = dev(w0), dev(v0), dev(gv0)
    ctx.sparse_wgrad_update(mode, -1.0 / batch, d_start, d_end, d_idx, d_data, dev(delta), hp["alpha"], hp["lam"],
                            hp["lam1"], hp["mu"], hp["mu1"], hp["t"], d_v, d_gv, d_w)
    ctx.sync()
    assert rel_err(host(d_w), w_ref) < TOL
    if mode != orc.SGD:
        assert rel_err(host(d_v), v_ref) < TOL
    if mode in (orc.ADADELTA, orc.ADAM):
        assert rel_err(host(d_gv), gv_ref) < TOL


# ------------------------------------------------------------------ a7-a9
ACTS = {"sigmoid": 0, "tanh": 1, "relu": 2, "linear": 3, "softmax": 7, "elu": 10, "lrelu": 11, "selu": 12}


@pytest.mark.parametrize("act", list(ACTS))
def test_activation(ctx, orc, dsb, act):
    rng = np.random.default_rng(11)
    z = (rng.standard_normal((64, 515)) * 3).astype(np.float32)
    ref = orc.activation(ACTS[act], z.copy(), 0.01, 1.6733, 1.0507)
    d = dev(z)
    ctx.activation(ACTS[act], d, 0.01, 1.6733, 1.0507)
    ctx.sync()
    assert rel_err(host(d), ref) < TOL


def _output_inputs(h, batch, stride, seed=21):
    rng = np.random.default_rng(seed)
    z = (rng.standard_normal((batch, stride)) * 2.0 - 1.0).astype(np.float32)
    return z


@pytest.mark.parametrize("ef,act,weighted,iz", [
    ("smce", "sigmoid", False, False), ("smce", "sigmoid", True, False), ("smce", "sigmoid", False, True),
    ("ce", "sigmoid", False, False), ("ce", "sigmoid", True, True),
    ("l2", "sigmoid", False, False), ("l2", "linear", True, False), ("l2", "relu", False, False),
    ("l2", "tanh", False, False), ("l2", "lrelu", False, True),
    ("ce", "softmax", False, False), ("ce", "softmax", True, False), ("smce", "softmax", False, False),
    ("l2", "softmax", False, False)])
def test_sparse_loss_and_delta(ctx, orc, dsb, ef, act, weighted, iz):
    EF = {"l2": 1, "ce": 2, "smce": 3}[ef]
    A = ACTS[act]
    h = tiny(examples=128, width=2050, weighted=weighted)                # 2050: rows not 16-byte aligned
    batch, stride = 128, 2050
    unit = orc.activation(A, _output_inputs(h, batch, stride), 0.01, 1.6733, 1.0507)
    smce = (0.9, 0.1, 1.0, 1.0) if ef != "smce" else (0.8, 0.05, 1.5, 0.75)
    boost = (2.0, 0.5) if ef != "smce" else (1.0, 1.0)
    params = orc.make_params(deltaBoost=boost, smce=smce)
    oc = to_oracle(orc, h)
    ref_loss = orc.sparse_loss(params, oc, EF, A, 0, batch, unit, iz)
    ref_delta = orc.sparse_output_delta(params, oc, EF, A, 0, batch, unit, np.zeros_like(unit), iz, 0.01, 1.6733, 1.0507)
    ctx.set_params(deltaBoost=boost, smce=smce)
    dd = to_device(dsb, h)
    d_unit = dev(unit)
    got_loss = ctx.sparse_loss(dd, EF, A, 0, batch, d_unit, iz)
    d_delta = dev(np.full_like(unit, 7.0))
    ctx.sparse_output_delta(dd, EF, A, 0, batch, d_unit, d_delta, iz, 0.01, 1.6733, 1.0507)
    ctx.sync()
    ctx.set_params()
    assert abs(got_loss - ref_loss) <= TOL * max(abs(ref_loss), 1.0)
    assert rel_err(host(d_delta), ref_delta) < TOL


@pytest.mark.parametrize("ef", ["smce", "ce", "l2"])
@pytest.mark.parametrize("want_unit", [False, True])
@pytest.mark.parametrize("kernel,stride", [("row", 27278), ("row", 40001), ("tile", 27278)])
def test_output_pass_fused(ctx, orc, dsb, ef, want_unit, kernel, stride):
    """kernel = "row": the one-pass bitmap kernel (Boolean targets); "tile": the two-phase tile kernel it falls back to.
    stride 40,001: three column segments per row, odd width (unaligned rows, scalar head / tail)."""
    import torch
    EF = {"l2": 1, "ce": 2, "smce": 3}[ef]
    h = ml20m(examples=64, width=stride)
    batch = 64
    z = _output_inputs(h, batch, stride)
    smce = (1.0, 0.0, 1.0, 1.0)                                        # samples/movielens/config.json
    params = orc.make_params(smce=smce)
    oc = to_oracle(orc, h)
    unit = orc.activation(orc.ACT_SIGMOID, z.copy())
    ref_loss = orc.sparse_loss(params, oc, EF, orc.ACT_SIGMOID, 0, batch, unit)
    ref_delta = orc.sparse_output_delta(params, oc, EF, orc.ACT_SIGMOID, 0, batch, unit, np.zeros_like(unit))
    ctx.set_params(smce=smce)
    d_z = dev(z)
    d_unit = torch.empty_like(d_z) if want_unit else None
    d_delta = torch.empty_like(d_z)
    acc = torch.zeros(1, dtype=torch.int64, device="cuda")
    ctx.set_option("output_tile_kernel", int(kernel == "tile"))
    try:
        ctx.output_pass(to_device(dsb, h), EF, dsb.ACT_SIGMOID, 0, batch, d_z, d_unit, d_delta, acc)
        ctx.sync()
    finally:
        ctx.set_option("output_tile_kernel", 0)
    ctx.set_params()
    got_loss = float(acc.item()) / float(1 << 30)
    assert abs(got_loss - ref_loss) <= TOL * max(abs(ref_loss), 1.0)
    assert rel_err(host(d_delta), ref_delta) < TOL
    if want_unit:
        assert rel_err(host(d_unit), unit) < TOL


# ------------------------------------------------------------------ a10
@pytest.mark.parametrize("act", ["sigmoid", "tanh", "relu", "lrelu", "elu", "selu", "linear"])
def test_hadamard(ctx, orc, dsb, act):
    rng = np.random.default_rng(31)
    unit = rng.standard_normal((33, 130)).astype(np.float32)
    if act == "sigmoid":
        unit = 1.0 / (1.0 + np.exp(-unit))
    delta = rng.standard_normal((33, 130)).astype(np.float32)
    ref = orc.hadamard(ACTS[act], unit, delta.copy(), 2.0, 0.01, 1.6733, 1.0507)
    d = dev(delta)
    ctx.hadamard(ACTS[act], dev(unit), d, 2.0, 0.01, 1.6733, 1.0507)
    ctx.sync()
    assert rel_err(host(d), ref) < TOL


def test_sparseness_penalty(ctx, orc, dsb):
    rng = np.random.default_rng(32)
    unit = rng.random((1024, 128)).astype(np.float32)
    delta = rng.standard_normal((1024, 128)).astype(np.float32)
    ref = orc.sparseness_penalty(unit, delta.copy(), 0.5, 2.0)
    d = dev(delta)
    ctx.sparseness_penalty(dev(unit), d, 0.5, 2.0)
    ctx.sync()
    assert rel_err(host(d), ref) < TOL


# ------------------------------------------------------------------ a11
@pytest.mark.parametrize("B,k,n", [(256, 128, 128), (64, 128, 2050), (1024, 128, 27278)])
def test_gemms_fp32(ctx, orc, dsb, B, k, n):
    rng = np.random.default_rng(41)
    A = rng.random((B, k)).astype(np.float32)
    W = (rng.standard_normal((k, n)) * 0.05).astype(np.float32)
    D = (rng.standard_normal((B, n)) * 0.1).astype(np.float32)
    C0 = rng.standard_normal((B, n)).astype(np.float32)
    ref_c = orc.gemm_fwd(A, W, C0.copy(), 1.0)
    ref_g = orc.gemm_dw(A, D, np.zeros((k, n), dtype=np.float32), -1.0 / B, 0.0)
    ref_x = orc.gemm_dx(D, W, np.zeros((B, k), dtype=np.float32), 0.0)
    ctx.set_option("gemm_mode", dsb.GEMM_FP32)
    dC, dG, dX = dev(C0), dev(np.zeros((k, n), dtype=np.float32)), dev(np.zeros((B, k), dtype=np.float32))
    ctx.gemm_fwd(dev(A), dev(W), dC, 1.0)
    ctx.gemm_dw(dev(A), dev(D), dG, -1.0 / B, 0.0)
    ctx.gemm_dx(dev(D), dev(W), dX, 0.0)
    ctx.sync()
    # fp32 GEMMs differ from the triple loop only by summation order
    assert rel_err(host(dC), ref_c) < 2e-5
    assert rel_err(host(dG), ref_g) < 2e-5
    assert rel_err(host(dX), ref_x) < 5e-5


# ------------------------------------------------------------------ a12
@pytest.mark.parametrize("mode", range(7))
def test_update_weights_and_biases(ctx, orc, dsb, mode):
    rng = np.random.default_rng(51 + mode)
    size = 128 * 1031
    w0 = (rng.standard_normal(size) * 0.05).astype(np.float32)
    g = (rng.standard_normal(size) * 0.01).astype(np.float32)
    v0 = (rng.random(size) * 0.01).astype(np.float32)
    gv0 = (rng.random(size) * 0.01).astype(np.float32)
    w_ref, v_ref, gv_ref = w0.copy(), v0.copy(), gv0.copy()
    orc.update_weights(mode, 0.025, 1e-4, 1e-5, 0.9, 0.999, 3.0, v_ref, g, gv_ref, w_ref)
    d_w, d_v, d_gv = dev(w0), dev(v0), dev(gv0)
    ctx.update_weights(mode, 0.025, 1e-4, 1e-5, 0.9, 0.999, 3.0, d_v, dev(g), d_gv, d_w)
    ctx.sync()
    assert rel_err(host(d_w), w_ref) < TOL
    for width in (128, 2050):
        delta = (rng.standard_normal((1024, width)) * 0.1).astype(np.float32)
        b0 = rng.standard_normal(width).astype(np.float32)
        bv0 = (rng.random(width) * 0.01).astype(np.float32)
        bgv0 = (rng.random(width) * 0.01).astype(np.float32)
        b_ref, bv_ref, bgv_ref = b0.copy(), bv0.copy(), bgv0.copy()
        orc.update_biases(mode, 0.025, 0.9, 0.999, 3.0, delta, bv_ref, bgv_ref, b_ref)
        d_b, d_bv, d_bgv = dev(b0), dev(bv0), dev(bgv0)
        ctx.update_biases(mode, 0.025, 0.9, 0.999, 3.0, dev(delta), d_bv, d_bgv, d_b)
        ctx.sync()
        # column means are summed in a different (tree) order than the oracle's serial loop
        assert rel_err(host(d_b), b_ref) < 5e-5


def test_regularization_error(ctx, orc, dsb):
    rng = np.random.default_rng(61)
    w = (rng.standard_normal(27278 * 128) * 0.05).astype(np.float32)
    ref = orc.regularization_error(1e-4, 1e-5, w)
    got = ctx.regularization_error(1e-4, 1e-5, dev(w))
    assert abs(got - ref) <= TOL * abs(ref)


# ------------------------------------------------------------------ a13
@pytest.mark.parametrize("B,K,N", [(128, 128, 1024), (128, 128, 100000), (128, 64, 1024), (128, 32, 64), (128, 1, 64),
                                   (64, 100, 27278), (16, 100, 1000000), (5, 256, 300), (3, 10, 7)])
def test_topk_shapes_of_reference_test(ctx, orc, dsb, B, K, N):
    """(B,K,N) of tst/gputests/TestSort.cpp:209-211 plus the BASELINE shapes; tie-free inputs."""
    import torch
    rng = np.random.default_rng(12345)
    scores = rng.permutation(B * N).astype(np.float32).reshape(B, N) if B * N < (1 << 24) else rng.random((B, N), dtype=np.float32)
    ref_k, ref_v = orc.topk(scores, K)
    ok = torch.empty((B, K), dtype=torch.float32, device="cuda")
    ov = torch.empty((B, K), dtype=torch.int32, device="cuda")
    ctx.topk(dev(scores), K, ok, ov)
    ctx.sync()
    np.testing.assert_array_equal(host(ok), ref_k)
    np.testing.assert_array_equal(u32(ov), ref_v)


def test_topk_ties_and_filter(ctx, orc, dsb):
    import torch
    rng = np.random.default_rng(77)
    B, N, K = 64, 50000, 100
    scores = rng.integers(0, 50, size=(B, N)).astype(np.float32)       # heavy ties: rule = lowest column first
    h = ml20m(examples=B, width=N)
    filt = (h.start, h.end, h.index)
    ref_k, ref_v = orc.topk(scores, K, filt=filt)
    dcsr = to_device(dsb, h)
    ok = torch.empty((B, K), dtype=torch.float32, device="cuda")
    ov = torch.empty((B, K), dtype=torch.int32, device="cuda")
    ctx.topk(dev(scores), K, ok, ov, filt=(dcsr.start, dcsr.end, dcsr.index))
    ctx.sync()
    np.testing.assert_array_equal(host(ok), ref_k)
    np.testing.assert_array_equal(u32(ov), ref_v)


def test_topk_kv_merge_variant(ctx, orc, dsb):
    import torch
    rng = np.random.default_rng(78)
    B, N, K = 32, 4000, 50
    key = rng.permutation(B * N).astype(np.float32).reshape(B, N)
    val = rng.integers(0, 1 << 30, size=(B, N)).astype(np.uint32)
    ref_k, ref_v = orc.topk(key, K, value=val)
    ok = torch.empty((B, K), dtype=torch.float32, device="cuda")
    ov = torch.empty((B, K), dtype=torch.int32, device="cuda")
    ctx.topk_kv(dev(key), torch.from_numpy(val.view(np.int32).copy()).cuda(), K, ok, ov)
    ctx.sync()
    np.testing.assert_array_equal(host(ok), ref_k)
    np.testing.assert_array_equal(u32(ov), ref_v)


@pytest.mark.parametrize("shards", [2, 8])
def test_topk_sharded_merge_on_one_gpu(ctx, orc, dsb, shards):
    """The model-parallel top-K scheme (NNNetwork::CalculateTopKGlobal) with the column shards of `shards` ranks taken one
    after the other on ONE GPU: local dsb200_topk, dsb200_topk_offset, rank-major concatenation, dsb200_topk_kv -- bit-exact
    against the oracle's top-K of the whole rows, heavy ties included (CPU restatement: tests/test_topk_global_scheme.py)."""
    import torch
    rng = np.random.default_rng(79)
    B, N, K = 32, 20011, 64
    scores = rng.integers(0, 40, size=(B, N)).astype(np.float32)
    want_k, want_v = orc.topk(scores, K)
    d = dev(scores)
    keys, vals = [], []
    for r in range(shards):
        lo, hi = orc.shard_range(N, r, shards)
        ok = torch.empty((B, K), dtype=torch.float32, device="cuda")
        ov = torch.empty((B, K), dtype=torch.int32, device="cuda")
        ctx.topk(d[:, lo:hi].contiguous(), K, ok, ov)
        ctx.topk_offset(ov, lo)
        keys.append(ok)
        vals.append(ov)
    fk, fv = torch.cat(keys, dim=1).contiguous(), torch.cat(vals, dim=1).contiguous()
    ok = torch.empty((B, K), dtype=torch.float32, device="cuda")
    ov = torch.empty((B, K), dtype=torch.int32, device="cuda")
    ctx.topk_kv(fk, fv, K, ok, ov)
    ctx.sync()
    np.testing.assert_array_equal(host(ok), want_k)
    np.testing.assert_array_equal(u32(ov), want_v)


@pytest.mark.parametrize("act", [0, 1, 2, 10, 12], ids=["sigmoid", "tanh", "relu", "elu", "selu"])
def test_dropout_matches_oracle_with_the_same_uniforms(ctx, orc, act):
    """dsb200_dropout draws its uniforms from the counter-based generator of dsb200_fill_uniform (restated in numpy in
    test_gpu_engine.fill_uniform_host); fed the same uniforms the oracle's NNLayer::CalculateDropout must agree exactly,
    also when the layer is split over column shards (mask independent of the sharding)."""
    import torch
    from test_gpu_engine import fill_uniform_host
    B, S, p, seed, stream = 64, 200, 0.35, 12134, 77
    rng = np.random.default_rng(act)
    u = rng.standard_normal((B, S)).astype(np.float32)
    rnd = fill_uniform_host(B * S, seed, stream).reshape(B, S)
    want = orc.dropout(act, u.copy(), rnd, p)
    d = torch.from_numpy(u.copy()).cuda()
    ctx.dropout(act, d, p, seed, stream)
    ctx.sync()
    got = d.cpu().numpy()
    assert rel_err(got, want) < 1e-6
    if act not in (10, 12):                                     # a * x + b of ELU / SELU may contract to an FMA on the device
        np.testing.assert_array_equal(got, want)
    assert 0.25 < float((rnd < p).mean()) < 0.45
    lo, hi = 50, 130                                            # a column shard [50, 130) of the same layer
    d2 = torch.from_numpy(np.ascontiguousarray(u[:, lo:hi])).cuda()
    ctx.dropout(act, d2, p, seed, stream, full_stride=S, col_offset=lo)
    ctx.sync()
    np.testing.assert_array_equal(d2.cpu().numpy(), got[:, lo:hi])
