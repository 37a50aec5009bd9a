"""Finite-difference check of the oracle network's back-propagation against its own forward loss -- the CPU re-host of
NNNetwork::Validate (E/NNNetwork.cpp:2131-2278: perturb a weight by delta, compare (E(w+d) - E(w)) / d with the
analytical gradient), on the sparse autoencoder shape of tst/test_data/validate_*.json.  It catches a backward pass
that is self-consistently wrong, which parity against golden kernel outputs alone cannot."""
import numpy as np
import pytest

from helpers import tiny


@pytest.mark.parametrize("error", ["L2", "CE", "SMCE"])
def test_backward_matches_finite_differences(orc, error):
    sizes, batch = [96, 12, 96], 16
    ef = {"L2": orc.ERR_L2, "CE": orc.ERR_CE, "SMCE": orc.ERR_SMCE}[error]
    h = tiny(examples=batch, width=sizes[0], mean=9.0)
    oc = orc.Csr(h.start, h.end, h.index)
    rng = np.random.default_rng(3)
    net = orc.Network(sizes, error=ef, mode=orc.SGD, max_batch=batch)
    for i in range(2):
        net.W(i)[:] = rng.standard_normal(net.W(i).shape).astype(np.float32) * 0.3
        net.b(i)[:] = rng.standard_normal(net.b(i).shape).astype(np.float32) * 0.1
    net.s.params = orc.make_params(smce=(0.9, 0.1, 1.0, 1.0))
    net.set_input(oc, batch)
    net.backward(oc, oc, 0, batch)
    grads = [net.dW(i).copy() for i in range(2)]            # dW = -(1/batch) dE/dW  (sgemm_alpha of E/NNLayer.cpp:2191)
    eps = 2e-2
    worst = 0.0
    for i in range(2):
        W = net.W(i)
        picks = rng.choice(W.size, size=12, replace=False)
        if i == 0:                                          # input weights: only rows of items present in the batch have gradient
            present = np.unique(h.index)
            picks = [int(r) * W.shape[1] + int(c) for r, c in zip(rng.choice(present, 12), rng.integers(0, W.shape[1], 12))]
        for flat in picks:
            r, c = divmod(int(flat), W.shape[1])
            w0 = W[r, c]
            W[r, c] = w0 + eps
            ep = net.loss(oc, oc, 0, batch)
            W[r, c] = w0 - eps
            em = net.loss(oc, oc, 0, batch)
            W[r, c] = w0
            numeric = (ep - em) / (2 * eps)
            analytic = -batch * float(grads[i][r, c])
            worst = max(worst, abs(numeric - analytic) / max(abs(numeric), abs(analytic), 1e-2))
    assert worst < 2e-2, worst
