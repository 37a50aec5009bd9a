"""GPU parity for the launcher VARIANTS of the hot path that the main kernel tests reach only through their default form
("implemented but never exercised"):
  * Indexed data sets (the example -> row indirection of E/kernels.cu:751, 2045-2108 and the Indexed launchers of kLoss.cu / kDelta.cu)
    through the transposed matrix, the loss, the output delta, the fused output pass and the fused forward of the output layer;
  * byte-valued analog data (`unsigned char` scaled by 1/256, `char` by 1/128: E/kernels.cu:923, 998) through sparse-Z, its
    denoised form and the transposed matrix + weight gradient;
  * top-K at the shape of BASELINE config 5 (4,096 x 1,000,000 scores, K = 100, exclusion filter), rows checked against the oracle.
Everything goes through the C ABI; integers and fixed-point sums must match exactly, fp32 within 1e-5 (helpers.rel_err)."""
import numpy as np
import pytest

from helpers import canon_columns, ml20m, rel_err, tiny, to_device, to_oracle

pytestmark = pytest.mark.gpu
TOL = 1e-5


def dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.detach().cpu().numpy()


def u32(t):
    return host(t).view(np.uint32)


def _indexed(n_rows, n_examples, seed):
    rng = np.random.default_rng(seed)
    return rng.integers(0, n_rows, size=n_examples).astype(np.uint32)


def _byte_valued(h, dtype, seed=5):
    rng = np.random.default_rng(seed)
    if dtype == np.uint8:
        h.data = rng.integers(1, 256, size=h.nnz).astype(np.uint8)
    else:
        h.data = rng.integers(-128, 128, size=h.nnz).astype(np.int8)
    return h


# ------------------------------------------------------------------ byte-valued analog data
@pytest.mark.parametrize("dtype", [np.uint8, np.int8], ids=["uchar", "char"])
@pytest.mark.parametrize("denoised", [False, True])
@pytest.mark.parametrize("staged", [0, 1], ids=["warp", "staged"])
def test_sparse_z_byte_valued(ctx, orc, dsb, dtype, denoised, staged):
    h = _byte_valued(tiny(examples=256, analog=True, weighted=True), dtype)
    rng = np.random.default_rng(12345)
    stride, batch = 128, 256
    W = (rng.standard_normal((h.width, stride)) * 0.05).astype(np.float32)
    Z0 = rng.standard_normal((batch, stride)).astype(np.float32)
    rnd = rng.random(h.nnz).astype(np.float32) if denoised else None
    p = 0.2 if denoised else 0.0
    ref = orc.sparse_z(orc.make_params(denoising_p=p), to_oracle(orc, h, random=rnd), 0, batch, W, Z0.copy(), 1.0, denoised)
    ctx.set_params(denoising_p=p)
    ctx.set_option("z_staged_kernel", staged)
    try:
        dZ = dev(Z0)
        ctx.sparse_z(to_device(dsb, h, random=rnd), 0, batch, dev(W), dZ, 1.0, denoised)
        ctx.sync()
    finally:
        ctx.set_option("z_staged_kernel", 0)
        ctx.set_params()
    assert rel_err(host(dZ), ref) < TOL


def _transpose(ctx, orc, dsb, h, batch, position=0, ex_index=None, denoised=False, p=0.0):
    import torch
    rng = np.random.default_rng(99)
    rnd = rng.random(h.nnz).astype(np.float32) if denoised else None
    oc = to_oracle(orc, h, random=rnd, ex_index=ex_index)
    tstart, cap = orc.transposed_capacity(oc, h.width, batch)
    params = orc.make_params(denoising_p=p)
    r_end, r_idx, r_data = orc.sparse_transpose(params, oc, position, batch, tstart, cap, denoised)
    ctx.set_params(denoising_p=p)
    ctx.set_option("transpose_sort", 1)
    d_start = torch.from_numpy(tstart.view(np.int32).copy()).cuda()
    d_end = torch.zeros(h.width, dtype=torch.int32, device="cuda")
    d_idx = torch.zeros(max(cap, 1), dtype=torch.int32, device="cuda")
    d_data = torch.zeros(max(cap, 1), dtype=torch.float32, device="cuda") if r_data is not None else None
    ctx.sparse_transpose(to_device(dsb, h, random=rnd, ex_index=ex_index), position, batch, h.width, d_start, d_end, d_idx, d_data, denoised)
    ctx.sync()
    ctx.set_params()
    return (tstart, r_end, r_idx, r_data), (d_start, d_end, d_idx, d_data)


@pytest.mark.parametrize("dtype", [np.uint8, np.int8], ids=["uchar", "char"])
def test_transposed_matrix_and_gradient_byte_valued(ctx, orc, dsb, dtype):
    h = _byte_valued(tiny(examples=256, analog=True, weighted=True), dtype)
    (tstart, r_end, r_idx, r_data), (d_start, d_end, d_idx, d_data) = _transpose(ctx, orc, dsb, h, 256)
    np.testing.assert_array_equal(u32(d_end), r_end)
    g_idx, g_data = u32(d_idx), host(d_data)
    for c in np.nonzero(r_end > tstart)[0]:
        s, e = tstart[c], r_end[c]
        np.testing.assert_array_equal(g_idx[s:e], r_idx[s:e])
        np.testing.assert_array_equal(g_data[s:e], r_data[s:e])               # value * (1/256 | 1/128) * weight: exact in fp32
    # ... and the gradient over it: fixed-point sums, bit exact
    rng = np.random.default_rng(3)
    n = 128
    delta = (rng.standard_normal((256, n)) * 0.1).astype(np.float32)
    ref = orc.sparse_wgrad(orc.make_params(), -1.0 / 256, 0.0, tstart, r_end, r_idx, r_data, delta, np.zeros((h.width, n), dtype=np.float32))
    dW = dev(np.zeros((h.width, n), dtype=np.float32))
    ctx.sparse_wgrad(-1.0 / 256, 0.0, d_start, d_end, d_idx, d_data, dev(delta), dW)
    ctx.sync()
    np.testing.assert_array_equal(host(dW), ref)


# ------------------------------------------------------------------ Indexed data sets
@pytest.mark.parametrize("denoised", [False, True])
def test_transposed_matrix_indexed(ctx, orc, dsb, denoised):
    """400 examples over 200 stored rows, a batch that starts in the middle: columns and counts bit exact (kCalculateIndexedSparse
    TransposedMatrix / ...DenoisedMatrix, E/kernels.cu:2045-2108, 2330-2420)."""
    h = tiny(examples=200, weighted=denoised)
    ex_index = _indexed(200, 400, seed=3)
    (tstart, r_end, r_idx, r_data), (_, d_end, d_idx, d_data) = _transpose(ctx, orc, dsb, h, 256, position=100, ex_index=ex_index,
                                                                           denoised=denoised, p=0.25 if denoised else 0.0)
    np.testing.assert_array_equal(u32(d_end), r_end)
    g_idx = u32(d_idx)
    if r_data is None:
        for c in np.nonzero(r_end > tstart)[0]:
            np.testing.assert_array_equal(g_idx[tstart[c]:r_end[c]], r_idx[tstart[c]:r_end[c]])
    else:
        # the same stored row may appear twice in a batch (two examples pointing at it): equal (row, value) pairs, any order
        got = canon_columns(tstart, r_end, g_idx, host(d_data))
        want = canon_columns(tstart, r_end, r_idx, r_data)
        for (gr, gd), (wr, wd) in zip(got, want):
            np.testing.assert_array_equal(gr, wr)
            np.testing.assert_array_equal(gd, wd)


def _output_inputs(h, batch, stride, seed=4):
    rng = np.random.default_rng(seed)
    return (rng.standard_normal((batch, stride)) * 1.5).astype(np.float32)


@pytest.mark.parametrize("ef,weighted", [("smce", False), ("smce", True), ("ce", False), ("l2", True)])
def test_sparse_loss_and_delta_indexed(ctx, orc, dsb, ef, weighted):
    """kCalculateIndexedSparse*Error / *OutputDelta (E/kLoss.cu, E/kDelta.cu Indexed launchers): example -> row through the index."""
    EF = {"l2": 1, "ce": 2, "smce": 3}[ef]
    h = tiny(examples=96, width=2050, weighted=weighted)
    ex_index = _indexed(96, 300, seed=8)
    batch, stride, position = 128, 2050, 64
    unit = orc.activation(orc.ACT_SIGMOID, _output_inputs(h, batch, stride))
    smce = (0.8, 0.05, 1.5, 0.75)
    params = orc.make_params(smce=smce)
    oc = to_oracle(orc, h, ex_index=ex_index)
    ref_loss = orc.sparse_loss(params, oc, EF, orc.ACT_SIGMOID, position, batch, unit, False)
    ref_delta = orc.sparse_output_delta(params, oc, EF, orc.ACT_SIGMOID, position, batch, unit, np.zeros_like(unit), False, 0.01, 1.6733, 1.0507)
    ctx.set_params(smce=smce)
    dd = to_device(dsb, h, ex_index=ex_index)
    d_unit = dev(unit)
    got_loss = ctx.sparse_loss(dd, EF, orc.ACT_SIGMOID, position, batch, d_unit, False)
    d_delta = dev(np.full_like(unit, 7.0))
    ctx.sparse_output_delta(dd, EF, orc.ACT_SIGMOID, position, batch, d_unit, d_delta, False, 0.01, 1.6733, 1.0507)
    ctx.sync()
    ctx.set_params()
    assert abs(got_loss - ref_loss) <= TOL * max(abs(ref_loss), 1.0)
    assert rel_err(host(d_delta), ref_delta) < TOL


@pytest.mark.parametrize("ef", ["smce", "ce", "l2"])
def test_fused_output_kernels_indexed(ctx, orc, dsb, ef):
    """The one-pass output kernel (dsb200_output_pass) and the fused forward of the output layer (dsb200_gemm_fwd_output_pass) on an
    Indexed, shuffled-position batch: loss, delta and the bias-gradient column sums against the oracle's separate passes."""
    import torch
    EF = {"l2": 1, "ce": 2, "smce": 3}[ef]
    B, k, n, position = 128, 64, 4100, 32
    h = ml20m(examples=100, width=n, mean=40.0)
    ex_index = _indexed(100, 256, seed=11)
    rng = np.random.default_rng(21)
    A = rng.random((B, k)).astype(np.float32)
    W = (rng.standard_normal((k, n)) * 0.2).astype(np.float32)
    bias = (rng.standard_normal(n) * 0.5 - 1.0).astype(np.float32)
    z = (A.astype(np.float64) @ W.astype(np.float64) + bias).astype(np.float32)
    smce = (1.0, 0.0, 1.0, 1.0)
    params = orc.make_params(smce=smce)
    oc = to_oracle(orc, h, ex_index=ex_index)
    unit = orc.activation(orc.ACT_SIGMOID, z.copy())
    ref_loss = orc.sparse_loss(params, oc, EF, orc.ACT_SIGMOID, position, B, unit)
    ref_delta = orc.sparse_output_delta(params, oc, EF, orc.ACT_SIGMOID, position, B, unit, np.zeros_like(unit))
    ctx.set_params(smce=smce)
    dd = to_device(dsb, h, ex_index=ex_index)
    # one-pass kernel over z
    d_delta = torch.empty((B, n), device="cuda")
    acc = torch.zeros(1, dtype=torch.int64, device="cuda")
    ctx.output_pass(dd, EF, dsb.ACT_SIGMOID, position, B, dev(z), None, d_delta, acc)
    ctx.sync()
    got = float(acc.item()) / float(1 << 30)
    assert abs(got - ref_loss) <= TOL * max(abs(ref_loss), 1.0)
    assert rel_err(host(d_delta), ref_delta) < TOL
    # fused forward (3xTF32 product: the stated bound of tests/test_gpu_gemm.py)
    ctx.set_option("gemm_mode", 2)
    try:
        d_delta2 = torch.full((B, n), float("nan"), device="cuda")
        acc2 = torch.zeros(1, dtype=torch.int64, device="cuda")
        parts = torch.full((4 * ((B + 127) // 128), n), float("nan"), device="cuda")
        n_parts = ctx.gemm_fwd_output_pass(dd, EF, dsb.ACT_SIGMOID, position, dev(A), dev(W), dev(bias), None, d_delta2, acc2, parts)
        ctx.sync()
    finally:
        ctx.set_option("gemm_mode", 0)
        ctx.set_params()
    got2 = float(acc2.item()) / float(1 << 30)
    assert abs(got2 - ref_loss) <= 3e-5 * max(abs(ref_loss), 1.0)
    assert rel_err(host(d_delta2), ref_delta) < 3e-5
    colsum = host(parts[:n_parts]).astype(np.float64).sum(axis=0)
    want = ref_delta.astype(np.float64).sum(axis=0)
    assert np.abs(colsum - want).max() <= 3e-5 * max(np.abs(want).max(), 1.0) + 1e-6


# ------------------------------------------------------------------ BASELINE config 5
def test_topk_at_the_config5_shape(ctx, orc, dsb):
    """4,096 x 1,000,000 scores (16 GB), K = 100, exclusion filter of ~144 entries per row: 24 rows spread over the batch are
    checked against the oracle bit for bit (keys, indices, fixed tie rule), and every row is checked for order and for the filter."""
    import torch
    B, N, K = 4096, 1_000_000, 100
    g = torch.Generator(device="cuda").manual_seed(5)
    scores = torch.rand((B, N), device="cuda", generator=g)
    h = ml20m(examples=B, width=N)
    dcsr = to_device(dsb, h)
    ok = torch.empty((B, K), dtype=torch.float32, device="cuda")
    ov = torch.empty((B, K), dtype=torch.int32, device="cuda")
    ctx.topk(scores, K, ok, ov, filt=(dcsr.start, dcsr.end, dcsr.index))
    ctx.sync()
    keys, vals = host(ok), u32(ov)
    assert (np.diff(keys, axis=1) <= 0).all()                                  # descending in every row
    rows = np.unique(np.concatenate([np.arange(0, B, 179), [B - 1]]))
    sub = host(scores[torch.from_numpy(rows).cuda()])
    start = np.zeros(len(rows), dtype=np.uint64)
    end = np.zeros(len(rows), dtype=np.uint64)
    idx, pos = [], 0
    for i, r in enumerate(rows):
        e = h.index[int(h.start[r]):int(h.end[r])]
        start[i] = pos
        pos += len(e)
        end[i] = pos
        idx.append(e)
        assert not np.intersect1d(vals[r], e).size or (keys[r][np.isin(vals[r], e)] == 0).all()   # filtered items score 0 (U/Filters.cpp:49-67)
    ref_k, ref_v = orc.topk(sub, K, filt=(start, end, np.concatenate(idx).astype(np.uint32)))
    np.testing.assert_array_equal(keys[rows], ref_k)
    np.testing.assert_array_equal(vals[rows], ref_v)


@pytest.mark.parametrize("N", [37, 100, 512, 513, 1025, 4100])
@pytest.mark.parametrize("K", [1, 20, 100])
def test_topk_small_and_odd_widths_with_filter(ctx, orc, dsb, N, K):
    """the widths of a model-parallel shard (2,050 / 4 = 512 | 513 columns, rows not 16-byte aligned) and widths below one chunk or
    below K, with the exclusion filter: keys and indices bit for bit"""
    import torch
    rng = np.random.default_rng(N * 1000 + K)
    B = 64
    scores = (1.0 / (1.0 + np.exp(-rng.standard_normal((B, N)) * 2.0))).astype(np.float32)
    h = tiny(examples=B, width=N, mean=max(1.0, N * 0.02))
    ref_k, ref_v = orc.topk(scores, K, filt=(h.start, h.end, h.index))
    dcsr = to_device(dsb, h)
    ok = torch.empty((B, K), dtype=torch.float32, device="cuda")
    ov = torch.empty((B, K), dtype=torch.int32, device="cuda")
    ctx.topk(dev(scores), K, ok, ov, filt=(dcsr.start, dcsr.end, dcsr.index))
    ctx.sync()
    np.testing.assert_array_equal(host(ok), ref_k)
    np.testing.assert_array_equal(u32(ov), ref_v)


@pytest.mark.parametrize("P,K", [(4, 20), (8, 100), (2, 7)])
def test_topk_kv_merge_of_rank_lists(ctx, orc, dsb, P, K):
    """the merge step of the model-parallel top-K: P sorted lists of K (key, global id) pairs per row, some lists padded with the
    (-MAX_VALUE, 0) of a short shard"""
    import torch
    rng = np.random.default_rng(P * 100 + K)
    B = 64
    key = np.sort(rng.random((B, P, K)).astype(np.float32), axis=2)[:, :, ::-1].copy()
    val = rng.integers(0, 1 << 20, size=(B, P, K)).astype(np.uint32)
    key[:, P - 1, K // 2:] = -999999999999999.0                            # a shard with fewer than K valid candidates
    val[:, P - 1, K // 2:] = 0
    key2, val2 = key.reshape(B, P * K), val.reshape(B, P * K)
    ref_k, ref_v = orc.topk(key2, K, value=val2)
    ok = torch.empty((B, K), dtype=torch.float32, device="cuda")
    ov = torch.empty((B, K), dtype=torch.int32, device="cuda")
    ctx.topk_kv(dev(key2), torch.from_numpy(val2.view(np.int32).copy()).cuda(), K, ok, ov)
    ctx.sync()
    np.testing.assert_array_equal(host(ok), ref_k)
    np.testing.assert_array_equal(u32(ov), ref_v)


# ------------------------------------------------------------------ capacity table of a streamed batch, built on the device
@pytest.mark.parametrize("width", [1, 100, 8192, 8193, 27278, 100003])
@pytest.mark.parametrize("misaligned", [False, True])
def test_transposed_capacity_table_on_the_device(ctx, orc, dsb, width, misaligned):
    """dsb200_transposed_capacity (the streaming caller's replacement of the host pass of E/NNTypes.cpp:1631-1735): tStart = exclusive
    prefix of the per-column counts rounded up to 32, total = the capacity; widths on both sides of the 8,192-column tile of the scan,
    tables at 16-byte and at 4-byte alignment; equal to the oracle's table when the dataset is one batch"""
    import torch
    from dsstne_b200 import datagen
    rows = 300
    h = datagen.make_csr(rows, width, min(20.5, width * 0.6), dist="binomial", col="zipf" if width > 1000 else "uniform")
    counts = np.bincount(h.index[int(h.start[0]):int(h.end[rows - 1])], minlength=width).astype(np.uint64)
    aligned = (counts + 31) // 32 * 32
    want = np.concatenate([[0], np.cumsum(aligned)[:-1]]).astype(np.uint32)
    off = 1 if misaligned else 0
    d_count = torch.full((width + 8,), 0x7fffffff, dtype=torch.int32, device="cuda")
    d_start = torch.full((width + 8,), -1, dtype=torch.int32, device="cuda")
    d_total = torch.zeros(1, dtype=torch.int32, device="cuda")
    ctx.transposed_capacity(to_device(dsb, h), rows, width, d_count[off:], d_start[off:], d_total)
    ctx.sync()
    got = d_start.cpu().numpy().view(np.uint32)
    np.testing.assert_array_equal(got[off:off + width], want)
    assert (got[:off] == 0xffffffff).all() and (got[off + width:] == 0xffffffff).all()        # nothing written outside the table
    np.testing.assert_array_equal(d_count.cpu().numpy().view(np.uint32)[off:off + width], counts.astype(np.uint32))
    assert int(d_total.item()) == int(aligned.sum())
    tstart, cap = orc.transposed_capacity(to_oracle(orc, h), width, rows)
    np.testing.assert_array_equal(tstart, want)
    assert cap == int(aligned.sum())
