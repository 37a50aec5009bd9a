"""GPU parity of the C++ host engine (NNNetwork / NNLayer / NNWeight / NNDataSet mirror, driven through
include/dsstne_b200_engine.h) against the whole-network CPU oracle (orc_net_*), on BASELINE.json
config 1 (2,048 -> 128 -> 2,048 sparse autoencoder, batch 256) and small multi-layer variants.
Tolerance: 1e-5 relative (helpers.rel_err) for losses, activations, deltas and updated weights.
"""
import os

import numpy as np
import pytest

from helpers import rel_err, tiny, to_oracle

pytestmark = pytest.mark.gpu
TOL = 1e-5
MODES = ["SGD", "Momentum", "AdaGrad", "Nesterov", "RMSProp", "AdaDelta", "Adam"]


@pytest.fixture(scope="module")
def eng(dsb):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from dsstne_b200 import engine
    engine.startup(0, 1, 0, None, seed=12134)
    engine.use_torch_stream(0)
    yield engine
    engine.shutdown()


def mix64(z):
    z = np.asarray(z, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xbf58476d1ce4e5b9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94d049bb133111eb)
        return z ^ (z >> np.uint64(31))


def fill_uniform_host(n, seed, stream):
    """numpy restatement of dsb200_fill_uniform (csrc/random.cu): U(0,1] from (seed, stream, i)."""
    with np.errstate(over="ignore"):
        key = mix64(np.uint64(seed) ^ mix64(np.uint64(stream) + np.uint64(0x9e3779b97f4a7c15)))
        r = mix64(key + np.arange(n, dtype=np.uint64) * np.uint64(0x9e3779b97f4a7c15))
    return ((r >> np.uint64(40)).astype(np.float32) + np.float32(1.0)) * np.float32(1.0 / 16777216.0)


def build_pair(eng, orc, sizes, h, batch, mode, error="ScaledMarginalCrossEntropy", smce=(1.0, 0.0, 1.0, 1.0), denoising_p=0.0,
               sparseness=None, fusion=True, p_dropout=0.0):
    from dsstne_b200 import datagen
    hidden = sizes[1:-1]
    ds_in = eng.Dataset.from_host_csr("gl_input", h)
    ds_out = eng.Dataset.from_host_csr("gl_output", h)
    net = eng.Network(eng.autoencoder_json(hidden, error=error, smce=smce, denoising_p=denoising_p, sparseness=sparseness, p_dropout=p_dropout),
                      batch, [ds_in, ds_out])
    net.set_training_mode(mode)
    net.set_fusion(fusion)
    Ws, bs = datagen.make_weights(sizes, scale=0.05)
    names = ["Input"] + [f"Hidden{i + 1}" for i in range(len(hidden))] + ["Output"]
    for i in range(len(sizes) - 1):
        bs[i][:] = np.random.default_rng(i).standard_normal(bs[i].shape).astype(np.float32) * 0.1
        net.set_weights(names[i], names[i + 1], Ws[i], bs[i])
    ef = {"ScaledMarginalCrossEntropy": orc.ERR_SMCE, "CrossEntropy": orc.ERR_CE, "L2": orc.ERR_L2}[error]
    onet = orc.Network(sizes, error=ef, mode=mode, max_batch=batch)
    for i in range(len(sizes) - 1):
        onet.W(i)[:] = Ws[i]
        onet.b(i)[:] = bs[i]
    onet.s.params = orc.make_params(denoising_p=denoising_p, smce=(smce[0], smce[1], smce[2], smce[3]))
    if sparseness is not None:
        onet.s.sparsenessPenalty_p, onet.s.sparsenessPenalty_beta = sparseness
        for l in range(1, len(sizes) - 1):
            onet.s.sparsePenalty[l] = 1
    onet.s.denoising = 1 if denoising_p > 0 else 0
    return net, onet, names, (ds_in, ds_out)


@pytest.mark.parametrize("mode", range(7), ids=MODES)
def test_config1_train_steps_match_oracle(eng, orc, mode):
    sizes, batch = [2048, 128, 2048], 256
    h = tiny(examples=512, width=2048)
    net, onet, names, _ = build_pair(eng, orc, sizes, h, batch, mode)
    oc = to_oracle(orc, h)
    onet.set_input(oc, batch)
    hp = dict(alpha=0.025, lam=1e-4, lam1=0.0, mu=0.5, mu1=0.999)
    for step, pos in enumerate([0, 256, 0]):
        got = net.train_step(pos, hp["alpha"], hp["lam"], hp["lam1"], hp["mu"], hp["mu1"])
        want, _ = onet.train_step(oc, oc, pos, batch, hp["alpha"], hp["lam"], hp["lam1"], hp["mu"], hp["mu1"])
        assert abs(got - want) <= TOL * abs(want), f"loss at step {step}"
    # AdaGrad / RMSProp / Adam divide by sqrt(v) clamped at 1e-9 (1e-8 for Adam): a weight whose gradient is at
    # rounding-noise level gets that noise multiplied by up to alpha * 3e4 (alpha * 1e8 for Adam), so two exact fp32
    # implementations that sum the dense GEMMs in a different order (cuBLAS vs the oracle's loops) legitimately differ
    # by more than 1e-5 after a few steps.  With identical gradients the kernels agree to 1e-5 for every mode
    # (test_gpu_kernels.py::test_update_weights_and_biases, ::test_sparse_wgrad_update_fused_equals_unfused).
    # the fused training pass does not store the output activations (only the delta is needed); reading them afterwards
    # must still give a = f(z) of the last forward pass (NNLayer::MaterializeUnits)
    assert rel_err(net.get_units("Output").reshape(batch, 2048), onet.unit(2, batch)) < TOL
    tol = {orc.ADAGRAD: 2e-4, orc.RMSPROP: 2e-4, orc.ADAM: 2e-3}.get(mode, TOL)
    for i in range(2):
        W, b = net.get_weights(names[i], names[i + 1])
        assert rel_err(W.reshape(onet.W(i).shape), onet.W(i)) < tol
        assert rel_err(b, onet.b(i)) < max(tol, 5e-5)          # column means are summed in tree order on the GPU
    net.close()


@pytest.mark.parametrize("gemm_mode,tol", [(0, TOL), (2, 3e-5)], ids=["fp32", "tf32x3"])
@pytest.mark.parametrize("error", ["ScaledMarginalCrossEntropy", "CrossEntropy", "L2"])
def test_three_hidden_layers_with_penalty(eng, orc, error, gemm_mode, tol):
    """gemm_mode 2 = the tcgen05 3xTF32 kernels (bound stated in tests/test_gpu_gemm.py); 0 = cuBLAS fp32."""
    sizes, batch = [2048, 128, 64, 128, 2048], 128
    h = tiny(examples=256, width=2048)
    net, onet, names, _ = build_pair(eng, orc, sizes, h, batch, orc.SGD, error=error, sparseness=(0.5, 2.0))
    net.set_gemm_mode(gemm_mode)
    eng.set_option("gemm_tc_min_work", 0)                           # these shapes are small: force the tensor-core kernel anyway
    oc = to_oracle(orc, h)
    onet.set_input(oc, batch)
    try:
        for pos in (0, 128):
            got = net.train_step(pos, 0.05)
            want, _ = onet.train_step(oc, oc, pos, batch, 0.05)
            assert abs(got - want) <= tol * abs(want)
        for i in range(len(sizes) - 1):
            W, b = net.get_weights(names[i], names[i + 1])
            assert rel_err(W.reshape(onet.W(i).shape), onet.W(i)) < tol
        # hidden activations and deltas of the last step
        for l in range(1, len(sizes) - 1):
            u = net.get_units(names[l]).reshape(batch, sizes[l])
            assert rel_err(u, onet.unit(l, batch)) < tol
    finally:
        net.set_gemm_mode(0)
        eng.set_option("gemm_tc_min_work", 2048)
        net.close()


@pytest.mark.parametrize("gemm_mode,tol", [(0, TOL), (2, 3e-5)], ids=["fp32", "tf32x3"])
def test_wide_hidden_layers_like_config4(eng, orc, gemm_mode, tol):
    """BASELINE config 4 in miniature: 1,024-wide hidden layers (sparse-Z / sparse gradient with an 8-block column loop,
    dense layers large enough for the tcgen05 kernels with their fused bias + activation epilogue)."""
    sizes, batch = [4096, 1024, 1024, 4096], 256
    h = tiny(examples=512, width=4096, mean=40.0)
    net, onet, names, _ = build_pair(eng, orc, sizes, h, batch, orc.MOMENTUM)
    net.set_gemm_mode(gemm_mode)
    oc = to_oracle(orc, h)
    onet.set_input(oc, batch)
    try:
        for pos in (0, 256):
            got = net.train_step(pos, 0.01, 1e-4, 0.0, 0.5, 0.0)
            want, _ = onet.train_step(oc, oc, pos, batch, 0.01, 1e-4, 0.0, 0.5, 0.0)
            assert abs(got - want) <= tol * abs(want)
        for i in range(3):
            W, b = net.get_weights(names[i], names[i + 1])
            assert rel_err(W.reshape(onet.W(i).shape), onet.W(i)) < tol
            assert rel_err(b, onet.b(i)) < max(tol, 5e-5)
    finally:
        net.set_gemm_mode(0)
        net.close()


@pytest.mark.parametrize("error", ["ScaledMarginalCrossEntropy", "CrossEntropy", "L2"])
@pytest.mark.parametrize("sizes", [[4096, 128, 128, 4096], [4099, 96, 128, 4099], [4096, 1024, 1024, 4096]], ids=["fused", "fused-ragged", "declined-wide-hidden"])
def test_output_gemm_fused_with_the_output_pass(eng, orc, error, sizes):
    """engine option "fuse_output_gemm" (default on): the output layer's forward GEMM is deferred into the loss / delta pass and runs
    as dsb200_gemm_fwd_output_pass (3xTF32 mode; the bias gradient comes out of the same kernel).  Hidden widths above 128 are
    declined by the kernel and take the two-call path.  Losses and weights against the oracle at the tensor-core bound, and the
    units a later reader gets (top-K after a training step) are materialised by re-running the layer."""
    batch = 256
    h = tiny(examples=512, width=sizes[0], mean=40.0)
    net, onet, names, _ = build_pair(eng, orc, sizes, h, batch, orc.MOMENTUM, error=error)
    net.set_gemm_mode(2)
    eng.set_option("fuse_output_gemm", 1)
    oc = to_oracle(orc, h)
    onet.set_input(oc, batch)
    tol = 3e-5
    try:
        for pos in (0, 256):
            got = net.train_step(pos, 0.01, 1e-4, 0.0, 0.5, 0.0)
            want, _ = onet.train_step(oc, oc, pos, batch, 0.01, 1e-4, 0.0, 0.5, 0.0)
            assert abs(got - want) <= tol * abs(want)
        units = np.asarray(net.get_units("Output")).reshape(batch, -1)        # MaterializeUnits after a deferred forward
        assert np.isfinite(units).all() and units.min() >= 0.0 and units.max() <= 1.0
        for i in range(3):
            W, b = net.get_weights(names[i], names[i + 1])
            assert rel_err(W.reshape(onet.W(i).shape), onet.W(i)) < tol
            assert rel_err(b, onet.b(i)) < max(tol, 5e-5)
    finally:
        net.set_gemm_mode(0)
        net.close()


@pytest.mark.parametrize("gemm_mode,tol", [(0, TOL), (2, 3e-5)], ids=["fp32", "tf32x3"])
def test_baseline_config2_exact_network(eng, orc, gemm_mode, tol):
    """The benchmarked network itself (BASELINE.json config 2, what bench.py times): 27,278 -> 128 -> 128 -> 128 -> 27,278,
    batch 1,024, synthetic CSR at ML-20M density (log-normal rows, Zipf columns), sigmoid / SMCE(1,0,1,1), SGD with the CLI's
    defaults.  Two steps: loss of each and every weight / bias afterwards against the oracle network.  In 3xTF32 mode this is
    the fused output-layer forward (csrc/gemm_stream.cu) plus the streamed dW / dX kernels."""
    from helpers import ml20m
    sizes, batch = [27278, 128, 128, 128, 27278], 1024
    h = ml20m(examples=2 * batch, width=sizes[0])
    net, onet, names, _ = build_pair(eng, orc, sizes, h, batch, orc.SGD)
    net.set_gemm_mode(gemm_mode)
    oc = to_oracle(orc, h)
    onet.set_input(oc, batch)
    try:
        for pos in (0, batch):
            got = net.train_step(pos, 0.025, 1e-4, 0.0, 0.5, 0.0)
            want, _ = onet.train_step(oc, oc, pos, batch, 0.025, 1e-4, 0.0, 0.5, 0.0)
            assert abs(got - want) <= tol * abs(want), (pos, got, want)
        for i in range(len(sizes) - 1):
            W, b = net.get_weights(names[i], names[i + 1])
            assert rel_err(W.reshape(onet.W(i).shape), onet.W(i)) < tol, names[i]
            assert rel_err(b, onet.b(i)) < max(tol, 5e-5), names[i]
    finally:
        net.set_gemm_mode(0)
        net.close()


@pytest.mark.parametrize("gemm_mode,tol", [(0, TOL), (2, 3e-5)], ids=["fp32", "tf32x3"])
def test_movielens_sample_configuration(eng, orc, gemm_mode, tol):
    """The shipped sample (samples/movielens/config.json of the reference): ONE 128-unit sigmoid hidden layer with the sparseness
    penalty, denoising p = 0.2 on the sparse input, ScaledMarginalCrossEntropy (1, 0, 1, 1), on ML-20M-shaped data.  The denoising
    randoms are generated on the device and handed to the oracle (cuRAND's stream is "parity unpinned")."""
    from helpers import ml20m
    sizes, batch = [27278, 128, 27278], 1024
    h = ml20m(examples=2 * batch, width=sizes[0])
    net, onet, names, _ = build_pair(eng, orc, sizes, h, batch, orc.SGD, denoising_p=0.2, sparseness=(0.5, 2.0))
    net.set_gemm_mode(gemm_mode)
    rnd = fill_uniform_host(h.nnz, 12134, 0)                     # what NNDataSet::GenerateDenoisingData draws for epoch 0 on rank 0
    oc = to_oracle(orc, h, random=rnd)
    onet.set_input(oc, batch)
    try:
        got = net.train(1, 0.025)                                # one epoch = two minibatches (Train regenerates the randoms per epoch)
        tot = 0.0
        for pos in (0, batch):
            e, _ = onet.train_step(oc, to_oracle(orc, h), pos, batch, 0.025)
            tot += e
        want = tot / (2 * batch)
        assert abs(got - want) <= tol * abs(want), (got, want)
        for i in range(len(sizes) - 1):
            W, b = net.get_weights(names[i], names[i + 1])
            assert rel_err(W.reshape(onet.W(i).shape), onet.W(i)) < tol, names[i]
            assert rel_err(b, onet.b(i)) < max(tol, 5e-5), names[i]
    finally:
        net.set_gemm_mode(0)
        net.close()


def test_fused_and_unfused_engine_agree(eng, orc):
    sizes, batch = [2048, 128, 2048], 256
    h = tiny(examples=256, width=2048)
    res = []
    for fusion in (True, False):
        net, _, names, _ = build_pair(eng, orc, sizes, h, batch, orc.MOMENTUM, fusion=fusion)
        losses = [net.train_step(0, 0.025, 1e-4, 0.0, 0.5, 0.0) for _ in range(3)]
        res.append((losses, [net.get_weights(names[i], names[i + 1]) for i in range(2)]))
        net.close()
    for a, b in zip(res[0][0], res[1][0]):
        assert abs(a - b) <= 1e-6 * abs(b)
    for (Wa, ba), (Wb, bb) in zip(res[0][1], res[1][1]):
        assert rel_err(Wa, Wb) < 1e-6
        assert rel_err(ba, bb) < 1e-6


def test_denoising_uses_the_counter_based_generator(eng, orc):
    """Train(1 epoch) with Denoising p=0.2: the engine draws its randoms from dsb200_fill_uniform; the oracle is
    fed the numpy restatement of that generator, so the epoch must agree end to end."""
    sizes, batch = [2048, 128, 2048], 256
    h = tiny(examples=512, width=2048)
    net, onet, names, _ = build_pair(eng, orc, sizes, h, batch, orc.SGD, denoising_p=0.2)
    rnd = fill_uniform_host(h.nnz, 12134, 0)
    oc = to_oracle(orc, h, random=rnd)
    onet.set_input(oc, batch)
    got = net.train(1, 0.025)
    tot = 0.0
    for pos in (0, 256):
        e, _ = onet.train_step(oc, to_oracle(orc, h), pos, batch, 0.025)
        tot += e
    want = tot / 512
    assert abs(got - want) <= TOL * abs(want)
    W, _ = net.get_weights(names[0], names[1])
    assert rel_err(W.reshape(onet.W(0).shape), onet.W(0)) < TOL
    net.close()


def test_dropout_training_matches_oracle(eng, orc):
    """pDropout 0.5 on every hidden layer (the reference's benchmark config, benchmarks/dsstne/config.json): the engine draws
    each mask inside dsb200_dropout from (seed, stream = 1<<62 | layer order << 40 | call number); the oracle network is fed
    the numpy restatement of the same uniforms.  Prediction must not drop anything."""
    sizes, batch = [2048, 128, 64, 2048], 128
    h = tiny(examples=256, width=2048)
    net, onet, names, _ = build_pair(eng, orc, sizes, h, batch, orc.SGD, p_dropout=0.5)
    oc = to_oracle(orc, h)
    onet.set_input(oc, batch)
    for step, pos in enumerate((0, 128, 0)):
        for l in (1, 2):
            stream = (1 << 62) | (l << 40) | step
            onet.set_dropout(l, 0.5, fill_uniform_host(batch * sizes[l], 12134, stream).reshape(batch, sizes[l]))
        got = net.train_step(pos, 0.05)
        want, _ = onet.train_step(oc, oc, pos, batch, 0.05)
        assert abs(got - want) <= TOL * abs(want), f"step {step}"
    for i in range(3):
        W, b = net.get_weights(names[i], names[i + 1])
        assert rel_err(W.reshape(onet.W(i).shape), onet.W(i)) < TOL
    net.set_position(0)
    net.predict_batch()
    onet.forward(oc, 0, batch)                                      # training=False: no dropout
    assert rel_err(net.get_units("Output").reshape(batch, 2048), onet.unit(3, batch)) < TOL
    net.close()


# engine option "pinned_mirror" (single host copy per streamed batch, the default) and the two-copy staging path behind it
STREAM_PATHS = [0, 1]


@pytest.mark.parametrize("pinned_mirror", STREAM_PATHS, ids=lambda v: "pinned-mirror" if v else "staging")
def test_streamed_batches_through_load_sparse_match_oracle(eng, orc, pinned_mirror):
    """The serving / streaming path bench.py's e2e number runs: a dataset the size of ONE batch is re-loaded every step with
    NNDataSet::LoadSparseData (pinned staging + asynchronous copies) and its transposed capacity table is rebuilt on the
    device (dsb200_transposed_capacity).  Same losses and weights as the oracle stepping through the resident dataset."""
    from dsstne_b200 import datagen
    sizes, batch = [2048, 128, 2048], 256
    h = tiny(examples=3 * batch, width=2048)
    parts = []
    for b in range(3):
        s0 = int(h.start[b * batch])
        parts.append(((h.start[b * batch:(b + 1) * batch] - np.uint64(s0)).astype(np.uint64), (h.end[b * batch:(b + 1) * batch] - np.uint64(s0)).astype(np.uint64),
                      np.ascontiguousarray(h.index[s0:int(h.end[(b + 1) * batch - 1])])))
    big = max(parts, key=lambda p: len(p[2]))
    ds_in = eng.Dataset("gl_input", big[0], big[1], big[2], 2048)
    ds_out = eng.Dataset("gl_output", big[0], big[1], big[2], 2048)
    net = eng.Network(eng.autoencoder_json([128]), batch, [ds_in, ds_out])
    net.set_training_mode(orc.MOMENTUM)
    Ws, bs = datagen.make_weights(sizes, scale=0.05)
    names = ["Input", "Hidden1", "Output"]
    for i in range(2):
        net.set_weights(names[i], names[i + 1], Ws[i], bs[i])
    onet = orc.Network(sizes, error=orc.ERR_SMCE, mode=orc.MOMENTUM, max_batch=batch)
    for i in range(2):
        onet.W(i)[:] = Ws[i]
        onet.b(i)[:] = bs[i]
    onet.s.params = orc.make_params(smce=(1.0, 0.0, 1.0, 1.0))
    oc = to_oracle(orc, h)
    onet.set_input(oc, batch)
    eng.set_option("pinned_mirror", pinned_mirror)
    try:
        for step, b in enumerate([0, 2, 1, 0]):
            st, en, ix = parts[b]
            ds_in.load_sparse(st, en, ix)
            ds_out.load_sparse(st, en, ix)
            got = net.train_step(0, 0.025, 1e-4, 0.0, 0.5, 0.0)
            want, _ = onet.train_step(oc, oc, b * batch, batch, 0.025, 1e-4, 0.0, 0.5, 0.0)
            assert abs(got - want) <= TOL * abs(want), f"step {step}"
    finally:
        eng.set_option("pinned_mirror", 1)
    for i in range(2):
        W, bb = net.get_weights(names[i], names[i + 1])
        assert rel_err(W.reshape(onet.W(i).shape), onet.W(i)) < TOL
        assert rel_err(bb, onet.b(i)) < 5e-5
    net.close()


def test_epoch_shuffle_is_a_permutation_and_trains_like_the_permuted_dataset(eng, dsb):
    """NNNetwork::ShuffleIndices (E/NNNetwork.cpp:874-907; the permutation itself is unpinned -- cuRAND in the reference, the counter
    based generator here): every epoch draws a fresh permutation of all examples from the identity, a function of (seed, epoch);
    an epoch with shuffling on gives the losses and weights of an epoch without shuffling over the data set laid out in that order"""
    from dsstne_b200 import datagen
    batch, examples, sizes = 64, 4 * 64, [2048, 64, 2048]
    h = tiny(examples=examples, width=2048)
    Ws, bs = datagen.make_weights(sizes, scale=0.05)
    names = ["Input", "Hidden1", "Output"]

    def network(data, shuffle, W, b):
        ds = [eng.Dataset.from_host_csr("gl_input", data), eng.Dataset.from_host_csr("gl_output", data)]
        net = eng.Network(eng.autoencoder_json([64], shuffle=shuffle), batch, ds)
        net.set_training_mode(dsb.SGD)
        for i in range(2):
            net.set_weights(names[i], names[i + 1], W[i], b[i])
        return net

    def weights(net):
        out = [net.get_weights(names[i], names[i + 1]) for i in range(2)]
        return [w for w, _ in out], [b for _, b in out]

    def permuted(perm):
        lens = (h.end - h.start).astype(np.uint64)[perm]
        end = np.cumsum(lens).astype(np.uint64)
        index = np.concatenate([h.index[int(h.start[i]):int(h.end[i])] for i in perm]).astype(np.uint32)
        return datagen.HostCsr(end - lens, end, index, h.width)

    a = network(h, True, Ws, bs)                                     # (one network at a time: the engine publishes ONE to the kernels)
    epochs = []
    for epoch in range(2):
        err = a.train(1, 0.025, 1e-4, 0.0, 0.0, 0.0)
        perm = a.get_shuffle_indices()
        assert perm.size == examples and np.array_equal(np.sort(perm), np.arange(examples, dtype=np.uint32))
        assert not np.array_equal(perm, np.arange(examples, dtype=np.uint32))
        epochs.append((err, perm, weights(a)))
    a.close()
    assert not np.array_equal(epochs[0][1], epochs[1][1])
    state = (Ws, bs)
    for epoch, (err_a, perm, after) in enumerate(epochs):
        b = network(permuted(perm), False, *state)                  # same start of the epoch, rows physically in the shuffled order
        err_b = b.train(1, 0.025, 1e-4, 0.0, 0.0, 0.0)
        Wb, bb = weights(b)
        b.close()
        assert abs(err_a - err_b) <= 1e-5 * abs(err_b), f"epoch {epoch}: {err_a} vs {err_b}"
        for i in range(2):
            assert rel_err(after[0][i], Wb[i]) < 1e-5 and rel_err(after[1][i], bb[i]) < 1e-5, f"epoch {epoch}, weight {i}"
        state = after
    c = network(h, True, Ws, bs)                                     # a new network draws the same sequence of permutations
    c.train(1, 0.025, 1e-4, 0.0, 0.0, 0.0)
    first = c.get_shuffle_indices()
    c.close()
    assert np.array_equal(first, epochs[0][1])


def test_predict_and_topk_with_filter(eng, orc):
    sizes, batch = [2048, 128, 2048], 256
    h = tiny(examples=256, width=2048)
    net, onet, names, (ds_in, _) = build_pair(eng, orc, sizes, h, batch, orc.SGD)
    oc = to_oracle(orc, h)
    net.set_position(0)
    net.predict_batch()
    onet.forward(oc, 0, batch)
    scores = net.get_units("Output").reshape(batch, 2048)
    assert rel_err(scores, onet.unit(2, batch)) < TOL
    key, val = net.topk("Output", 100, batch, filt=ds_in)
    want_k, want_v = orc.topk(scores, 100, filt=(h.start, h.end, h.index))     # same scores: selection must be exact
    np.testing.assert_array_equal(key, want_k)
    np.testing.assert_array_equal(val, want_v)
    net.close()


def test_netcdf_datasets_and_network_checkpoint_round_trip(eng, orc, tmp_path):
    """SaveNetCDF -> LoadNetCDF of the datasets and NNNetwork::SaveNetCDF -> LoadNeuralNetworkNetCDF of the trained
    network reproduce the same training step and the same predictions (checkpoint / resume, E/NNNetwork.cpp:1678-1691)."""
    sizes, batch = [2048, 128, 2048], 256
    h = tiny(examples=512, width=2048)
    net, onet, names, (ds_in, ds_out) = build_pair(eng, orc, sizes, h, batch, orc.MOMENTUM)
    for pos in (0, 256):
        net.train_step(pos, 0.025, 1e-4, 0.0, 0.5, 0.0)
    data_nc, net_nc = str(tmp_path / "data.nc"), str(tmp_path / "net.nc")
    eng.save_netcdf(data_nc, [ds_in, ds_out])
    net.save_netcdf(net_nc)
    loaded = eng.load_netcdf(data_nc)
    assert [d.name for d in loaded] == ["gl_input", "gl_output"] and loaded[0].examples == 512 and loaded[0].width == 2048
    assert loaded[0].nnz == h.nnz and loaded[0].attributes & 3 == 3                     # Sparse | Boolean
    net2 = eng.Network.from_netcdf(net_nc, batch, loaded)
    net2.set_training_mode(orc.MOMENTUM)
    for i in range(2):
        Wa, ba = net.get_weights(names[i], names[i + 1])
        Wb, bb = net2.get_weights(names[i], names[i + 1])
        np.testing.assert_array_equal(Wa, Wb)
        np.testing.assert_array_equal(ba, bb)
    net.set_position(0); net.predict_batch()
    net2.set_position(0); net2.predict_batch()
    np.testing.assert_array_equal(net.get_units("Output"), net2.get_units("Output"))
    # the optimizer state is not part of a DSSTNE checkpoint (weights + biases only): compare a velocity-free step
    net.set_training_mode(orc.SGD); net2.set_training_mode(orc.SGD)
    a = net.train_step(256, 0.025, 1e-4)
    b = net2.train_step(256, 0.025, 1e-4)
    assert a == b
    net.close(); net2.close()


@pytest.mark.parametrize("flavour", ["old", "new"])
def test_netcdf4_dataset_file_loads_and_trains(eng, orc, flavour):
    """a netCDF-4 (HDF5) dataset file in the schema of U/NetCDFhelper.cpp:332-372 -- the committed fixtures of
    tests/golden/make_hdf5_fixture.py -- through LoadNetCDF, then one training step against the oracle on the same data"""
    import importlib.util
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_hdf5_fixture", os.path.join(here, "golden", "make_hdf5_fixture.py"))
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
    start, end, index, data, gatts = m.sample()
    loaded = eng.load_netcdf(os.path.join(here, "golden", f"dataset_nc4_{flavour}.nc"))
    assert [d.name for d in loaded] == ["gl_input"] and loaded[0].examples == 5 and loaded[0].width == 256 and loaded[0].nnz == 12
    assert loaded[0].attributes & 3 == 1                                                # Sparse, analog (float values)
    out = eng.Dataset("gl_output", start.astype(np.uint64), end.astype(np.uint64), index, 256)      # Boolean targets
    sizes, batch = [256, 16, 256], 5
    net = eng.Network(eng.autoencoder_json([16]), batch, [loaded[0], out])
    net.set_training_mode(orc.SGD)
    from dsstne_b200 import datagen
    Ws, bs = datagen.make_weights(sizes, scale=0.1)
    net.set_weights("Input", "Hidden1", Ws[0], bs[0])
    net.set_weights("Hidden1", "Output", Ws[1], bs[1])
    onet = orc.Network(sizes, error=orc.ERR_SMCE, mode=orc.SGD, max_batch=batch)
    for i in range(2):
        onet.W(i)[:] = Ws[i]; onet.b(i)[:] = bs[i]
    onet.s.params = orc.make_params(smce=(1.0, 0.0, 1.0, 1.0))
    oc_in = orc.Csr(start.astype(np.uint64), end.astype(np.uint64), index, data=data)
    oc_out = orc.Csr(start.astype(np.uint64), end.astype(np.uint64), index)
    onet.set_input(oc_in, batch)
    got = net.train_step(0, 0.05, 0.0)
    want, _ = onet.train_step(oc_in, oc_out, 0, batch, 0.05, 0.0)
    assert abs(got - want) <= TOL * abs(want), (got, want)
    W, _ = net.get_weights("Input", "Hidden1")
    assert rel_err(W.reshape(onet.W(0).shape), onet.W(0)) < TOL
    net.close()


def test_engine_rejects_features_outside_the_hot_path(eng):
    from dsstne_b200 import DsbError
    h = tiny(examples=64, width=256)
    ds = eng.Dataset.from_host_csr("gl_input", h)
    bad = '{"Version":0.8,"Layers":[{"Kind":"Input","N":"auto","DataSet":"gl_input","Sparse":true},' \
          '{"Kind":"Hidden","Type":"Convolutional","N":16},{"Kind":"Output","N":"auto","DataSet":"gl_input","Sparse":true}]}'
    with pytest.raises(DsbError):
        eng.Network(bad, 32, [ds])
    with pytest.raises(DsbError):
        eng.Network('{"Version":0.8,"Bogus":1,"Layers":[]}', 32, [ds])        # unknown key is fatal, as in the reference


@pytest.mark.parametrize("error,smce", [("L2", None), ("ScaledMarginalCrossEntropy", (1.0, 0.0, 30.0, 1.0)), ("CrossEntropy", None)])
@pytest.mark.parametrize("hidden", [[8], [12, 8]], ids=["one-hidden", "two-hidden"])
def test_validate_finite_differences_on_the_gpu_engine(eng, orc, error, smce, hidden):
    """NNNetwork::Validate (E/NNNetwork.cpp:2459-2633) run on the product: the networks of tst/test_data/validate_{L2,ScaledMarginalCrossEntropy}
    _0{1,2}.json (sparse input -> sigmoid hidden layer(s) -> sparse sigmoid output, oneScale 30 for SMCE) at a size where the sparse kernels
    do real work; every weight and bias gradient the training kernels produce must agree with a finite difference of the loss within 20e-3."""
    from dsstne_b200 import datagen
    width, batch = 40, 8
    h = tiny(examples=batch, width=width, mean=5.0)
    ds_in = eng.Dataset.from_host_csr("gl_input", h)
    ds_out = eng.Dataset.from_host_csr("gl_output", h)
    kw = {} if smce is None else {"smce": smce}
    net = eng.Network(eng.autoencoder_json(hidden, error=error, **kw), batch, [ds_in, ds_out])
    sizes = [width] + hidden + [width]
    Ws, bs = datagen.make_weights(sizes, scale=0.3)
    names = ["Input"] + [f"Hidden{i + 1}" for i in range(len(hidden))] + ["Output"]
    for i in range(len(sizes) - 1):
        net.set_weights(names[i], names[i + 1], Ws[i], bs[i])
    try:
        assert net.validate(samples=48) is True
        # the network is usable afterwards, with its fusions back on
        assert np.isfinite(net.train_step(0, 0.01))
    finally:
        net.close()
