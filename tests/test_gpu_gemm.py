"""Dense GEMMs (hot-path row a11) through the C ABI, all three arithmetic modes, against a float64 product.

Stated bounds (helpers.rel_err = max |a-b| / (|b| + rms(b))):
  DSB200_GEMM_FP32    cuBLAS SGEMM                                   1e-5
  DSB200_GEMM_TF32X3  tcgen05, 3xTF32 split, fp32 accumulation in TMEM  3e-5   (products are fp32-grade, ~2e-6 for K <= 128;
                      the tensor core's fp32 accumulator truncates, so the bound grows with the length of one
                      accumulation chain: 1.8e-5 measured at K = 1,024, 1e-5 at K = 27,278 split 37 ways)
  DSB200_GEMM_TF32    tcgen05, one tf32 MMA per k-step                  3e-3
Shapes: BASELINE.json config 2's output layer (1,024 x 128 x 27,278), its hidden layers, and ragged / odd sizes that
exercise zero fill, 8- and 4-byte copy paths and partial tiles."""
import os

import numpy as np
import pytest
import torch

from helpers import rel_err

pytestmark = pytest.mark.gpu
BOUND = {0: 1e-5, 2: 3e-5, 1: 3e-3}
SHAPES = [(1024, 128, 27278), (1024, 128, 128), (256, 128, 256), (300, 70, 1000), (130, 33, 259), (64, 1024, 2000), (1, 5, 7),
          (200, 36, 516), (257, 129, 131)]   # 16-byte aligned rows with ragged K / N; every dimension odd


def ref64(a):
    return a.double().cpu().numpy()


LOADERS = [(0, -1), (2, -1), (2, 2), (2, 1), (2, 0), (1, 2), (1, 1), (1, 0)]
LOADER_IDS = ["fp32", "tf32x3-auto", "tf32x3-tmemA", "tf32x3-regload", "tf32x3-cpasync", "tf32-tmemA", "tf32-regload", "tf32-cpasync"]
# gemm_loader = 3: coalesced tensor-memory A loader on tcgen05.st.16x256b (K-major A)
LOADERS += [(2, 3), (1, 3)]
LOADER_IDS += ["tf32x3-tmemA16x256", "tf32-tmemA16x256"]


@pytest.mark.parametrize("mode,loader", LOADERS, ids=LOADER_IDS)
@pytest.mark.parametrize("B,k,n", SHAPES)
def test_gemm_fwd_dw_dx(ctx, mode, loader, B, k, n):
    g = torch.Generator(device="cuda").manual_seed(B * 7 + k * 3 + n)
    A = torch.randn(B, k, device="cuda", generator=g)
    W = torch.randn(k, n, device="cuda", generator=g) * 0.1
    D = torch.randn(B, n, device="cuda", generator=g) * 0.1
    C0 = torch.randn(B, n, device="cuda", generator=g)
    ctx.set_option("gemm_mode", mode)
    ctx.set_option("gemm_tc_min_work", 0)                       # force the tensor-core kernel for every shape
    ctx.set_option("gemm_loader", loader)                       # operand path: -1 per shape, 2 A through tensor memory, 1 registers, 0 cp.async + split warps
    try:
        C = C0.clone()
        ctx.gemm_fwd(A, W, C, beta=1.0)
        G = torch.zeros(k, n, device="cuda")
        ctx.gemm_dw(A, D, G, -1.0 / B)
        G2 = G.clone()
        ctx.gemm_dw(A, D, G2, -1.0 / B, beta=1.0)
        Dp = torch.zeros(B, k, device="cuda")
        ctx.gemm_dx(D, W, Dp)
        ctx.sync()
    finally:
        ctx.set_option("gemm_mode", 0)
        ctx.set_option("gemm_tc_min_work", 2048)
        ctx.set_option("gemm_loader", -1)
    a, w, d = ref64(A), ref64(W), ref64(D)
    tol = BOUND[mode]
    assert rel_err(C.cpu().numpy(), ref64(C0) + a @ w) < tol
    assert rel_err(G.cpu().numpy(), (-1.0 / B) * (a.T @ d)) < tol
    assert rel_err(G2.cpu().numpy(), 2 * (-1.0 / B) * (a.T @ d)) < tol
    assert rel_err(Dp.cpu().numpy(), d @ w.T) < tol


@pytest.mark.parametrize("mode", [0, 2], ids=["fp32", "tf32x3"])
@pytest.mark.parametrize("act", [0, 1, 2, 3], ids=["sigmoid", "tanh", "relu", "linear"])
def test_fused_bias_activation_forward(ctx, orc, mode, act):
    B, k, n = 512, 128, 1000
    g = torch.Generator(device="cuda").manual_seed(act)
    A = torch.randn(B, k, device="cuda", generator=g)
    W = torch.randn(k, n, device="cuda", generator=g) * 0.1
    bias = torch.randn(n, device="cuda", generator=g)
    C = torch.empty(B, n, device="cuda")
    ctx.set_option("gemm_mode", mode)
    ctx.set_option("gemm_tc_min_work", 0)
    try:
        ctx.gemm_fwd_bias_act(A, W, bias, act, C)
        ctx.sync()
    finally:
        ctx.set_option("gemm_mode", 0)
        ctx.set_option("gemm_tc_min_work", 2048)
    z = (ref64(A) @ ref64(W) + ref64(bias)[None, :]).astype(np.float32)
    want = orc.activation(act, np.ascontiguousarray(z))              # E/kActivation.cu semantics via the oracle
    assert rel_err(C.cpu().numpy(), want) < BOUND[mode]


@pytest.mark.parametrize("act", [0, 1, 2, 3, 10, 11, 12], ids=["sigmoid", "tanh", "relu", "linear", "elu", "lrelu", "selu"])
@pytest.mark.parametrize("B,k,n", [(1024, 128, 128), (100, 70, 33), (256, 64, 200)])
def test_small_dense_fused_kernels(ctx, orc, act, B, k, n):
    """csrc/dense_small.cu: forward bias + GEMM + activation and input delta + Hadamard product, one SIMT launch each, against
    the oracle's separate steps (exact fp32: 1e-5)."""
    g = torch.Generator(device="cuda").manual_seed(B + act)
    A = torch.randn(B, k, device="cuda", generator=g)
    W = torch.randn(k, n, device="cuda", generator=g) * 0.1
    bias = torch.randn(n, device="cuda", generator=g)
    D = torch.randn(B, n, device="cuda", generator=g) * 0.1
    C = torch.empty(B, n, device="cuda")
    ctx.gemm_fwd_bias_act(A, W, bias, act, C, 0.01, 1.6733, 1.0507)
    unit = torch.rand(B, k, device="cuda", generator=g) * 2 - 0.5                # activation values of the layer below
    Dp = torch.empty(B, k, device="cuda")
    ctx.gemm_dx_hadamard(D, W, act, unit, Dp, 2.0, 0.01, 1.6733, 1.0507)
    ctx.sync()
    z = (ref64(A) @ ref64(W) + ref64(bias)[None, :]).astype(np.float32)
    assert rel_err(C.cpu().numpy(), orc.activation(act, np.ascontiguousarray(z), 0.01, 1.6733, 1.0507)) < 1e-5
    dx = (ref64(D) @ ref64(W).T).astype(np.float32)
    want = orc.hadamard(act, unit.cpu().numpy(), np.ascontiguousarray(dx), 2.0, 0.01, 1.6733, 1.0507)
    assert rel_err(Dp.cpu().numpy(), want) < 1e-5


def test_tensor_core_gemm_is_deterministic(ctx):
    B, k, n = 1024, 128, 27278
    g = torch.Generator(device="cuda").manual_seed(5)
    D = torch.randn(B, n, device="cuda", generator=g) * 0.1
    W = torch.randn(k, n, device="cuda", generator=g) * 0.1
    ctx.set_option("gemm_mode", 2)
    try:
        outs = []
        for _ in range(3):
            Dp = torch.zeros(B, k, device="cuda")
            ctx.gemm_dx(D, W, Dp)                                   # split-K: partial tiles summed in a fixed order
            ctx.sync()
            outs.append(Dp.cpu().numpy())
    finally:
        ctx.set_option("gemm_mode", 0)
    np.testing.assert_array_equal(outs[0], outs[1])
    np.testing.assert_array_equal(outs[0], outs[2])


@pytest.mark.parametrize("ef", [3, 2, 1], ids=["smce", "ce", "l2"])
@pytest.mark.parametrize("want_unit", [False, True])
@pytest.mark.parametrize("B,k,n", [(1024, 128, 27278), (256, 128, 4099), (192, 100, 1000), (992, 64, 333)])
def test_forward_gemm_with_fused_output_pass(ctx, dsb, ef, want_unit, B, k, n):
    """dsb200_gemm_fwd_output_pass (csrc/gemm_stream.cu) against the two calls it replaces (dsb200_gemm_fwd_bias_act with the
    activation deferred, then dsb200_output_pass), both in 3xTF32: z agrees to the tensor-core bound (the two kernels sum the
    products in different orders), so delta / activations within 3e-5 of each other and the loss within 1e-5 relative.  Also the
    column sums of delta (the bias gradient) and dsb200_update_biases_partials against dsb200_update_biases."""
    from helpers import ml20m, to_device
    g = torch.Generator(device="cuda").manual_seed(ef * 10 + want_unit)
    A = torch.rand(B, k, device="cuda", generator=g)
    W = torch.randn(k, n, device="cuda", generator=g) * 0.1
    bias = torch.randn(n, device="cuda", generator=g) * 0.5 - 2.0
    h = ml20m(examples=B, width=n, mean=min(144.4, n / 8))
    ds = to_device(dsb, h)
    ctx.set_params(smce=(1.0, 0.0, 30.0, 1.0))
    ctx.set_option("gemm_mode", 2)
    ctx.set_option("gemm_tc_min_work", 0)
    try:
        z = torch.empty(B, n, device="cuda")
        ctx.gemm_fwd_bias_act(A, W, bias, 3, z)                          # 3 = Linear: activation deferred to the output pass
        unit0 = torch.empty_like(z) if want_unit else None
        delta0 = torch.empty_like(z)
        acc0 = torch.zeros(1, dtype=torch.int64, device="cuda")
        ctx.output_pass(ds, ef, dsb.ACT_SIGMOID, 0, B, z, unit0, delta0, acc0)
        unit1 = torch.empty_like(z) if want_unit else None
        delta1 = torch.full_like(z, float("nan"))
        acc1 = torch.zeros(1, dtype=torch.int64, device="cuda")
        parts = torch.full((4 * ((B + 127) // 128), n), float("nan"), device="cuda")
        n_parts = ctx.gemm_fwd_output_pass(ds, ef, dsb.ACT_SIGMOID, 0, A, W, bias, unit1, delta1, acc1, parts)
        b0 = bias.clone(); b1 = bias.clone()
        v0 = torch.zeros(n, device="cuda"); v1 = torch.zeros(n, device="cuda")
        ctx.update_biases(dsb.MOMENTUM, 0.05, 0.5, 0.0, 0.0, delta1, v0, None, b0)
        ctx.update_biases_partials(dsb.MOMENTUM, 0.05, 0.5, 0.0, 0.0, B, parts, n_parts, v1, None, b1)
        ctx.sync()
    finally:
        ctx.set_option("gemm_mode", 0)
        ctx.set_option("gemm_tc_min_work", 2048)
        ctx.set_params()
    assert n_parts >= 2
    assert rel_err(delta1.cpu().numpy(), delta0.cpu().numpy()) < 3e-5
    if want_unit:
        assert rel_err(unit1.cpu().numpy(), unit0.cpu().numpy()) < 3e-5
    l0, l1 = float(acc0.item()), float(acc1.item())
    assert abs(l1 - l0) <= 1e-5 * abs(l0)
    sums = parts[:n_parts].sum(dim=0).cpu().numpy()
    assert rel_err(sums, delta1.double().sum(dim=0).cpu().numpy()) < 1e-5
    assert rel_err(b1.cpu().numpy(), b0.cpu().numpy()) < 1e-5
    assert rel_err(v1.cpu().numpy(), v0.cpu().numpy()) < 1e-5


def test_fused_output_pass_declines_what_it_does_not_cover(ctx, dsb):
    """hidden width above 128 or the exact-fp32 mode: DSB200_EUNSUPPORTED, the caller makes the two calls"""
    from helpers import ml20m, to_device
    B, k, n = 128, 256, 1000
    A = torch.rand(B, k, device="cuda"); W = torch.randn(k, n, device="cuda"); bias = torch.zeros(n, device="cuda")
    ds = to_device(dsb, ml20m(examples=B, width=n, mean=20))
    delta = torch.empty(B, n, device="cuda")
    ctx.set_option("gemm_mode", 2)
    try:
        with pytest.raises(dsb.DsbError):
            ctx.gemm_fwd_output_pass(ds, 3, dsb.ACT_SIGMOID, 0, A, W, bias, None, delta)
    finally:
        ctx.set_option("gemm_mode", 0)
    with pytest.raises(dsb.DsbError):
        ctx.gemm_fwd_output_pass(ds, 3, dsb.ACT_SIGMOID, 0, A[:, :128].contiguous(), W[:128].contiguous(), bias, None, delta)


@pytest.mark.parametrize("mode", [2, 1], ids=["tf32x3", "tf32"])
@pytest.mark.parametrize("B,k,n", [(1024, 128, 27278), (1000, 128, 5001), (333, 100, 2049), (64, 256, 700), (1024, 17, 999), (32, 130, 513)])
def test_streamed_dw_dx_against_general_kernel_and_fp64(ctx, mode, B, k, n):
    """csrc/gemm_stream.cu (option gemm_stream = 1, the default for narrow layers) vs float64, ragged shapes, alpha / beta, determinism"""
    g = torch.Generator(device="cuda").manual_seed(B + k + n)
    A = torch.randn(B, k, device="cuda", generator=g)
    W = torch.randn(k, n, device="cuda", generator=g) * 0.1
    D = torch.randn(B, n, device="cuda", generator=g) * 0.1
    ctx.set_option("gemm_mode", mode)
    ctx.set_option("gemm_tc_min_work", 0)
    try:
        outs = []
        for stream in (1, 1, 0):
            ctx.set_option("gemm_stream", stream)
            G = torch.full((k, n), float("nan"), device="cuda")
            ctx.gemm_dw(A, D, G, -1.0 / B)
            G2 = G.clone()
            ctx.gemm_dw(A, D, G2, -1.0 / B, beta=1.0)
            Dp = torch.full((B, k), float("nan"), device="cuda")
            ctx.gemm_dx(D, W, Dp)
            Dp2 = Dp.clone()
            ctx.gemm_dx(D, W, Dp2, beta=1.0)
            ctx.sync()
            outs.append((G.cpu().numpy(), G2.cpu().numpy(), Dp.cpu().numpy(), Dp2.cpu().numpy()))
    finally:
        ctx.set_option("gemm_mode", 0)
        ctx.set_option("gemm_tc_min_work", 2048)
        ctx.set_option("gemm_stream", 1)
    a, w, d = ref64(A), ref64(W), ref64(D)
    tol = BOUND[mode]
    for G, G2, Dp, Dp2 in outs:
        assert rel_err(G, (-1.0 / B) * (a.T @ d)) < tol
        assert rel_err(G2, 2 * (-1.0 / B) * (a.T @ d)) < tol
        assert rel_err(Dp, d @ w.T) < tol
        assert rel_err(Dp2, 2 * (d @ w.T)) < tol
    for x, y in zip(outs[0], outs[1]):
        np.testing.assert_array_equal(x, y)                               # run to run
