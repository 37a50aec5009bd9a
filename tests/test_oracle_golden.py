"""Pins the CPU oracle (oracle/dsstne_oracle.c) against outputs of the REFERENCE itself:
  * tests/golden/ref_kernels.npz -- the reference's own CUDA kernels (kernels.cu, kLoss.cu, kDelta.cu, kActivation.cu,
    compiled unmodified for sm_100, oracle/Makefile) run on a B200 by tests/golden/make_ref_golden.py;
  * tests/golden/ref_topksort.npz -- the reference's CPU top-K comparator topKsort (U/Utils.cpp:213-243) run by
    tests/golden/make_topksort_golden.py, on the shapes of tst/gputests/TestSort.cpp:209-211.
The reference builds with -use_fast_math (Makefile.inc:64): exp/log/division outputs get 2e-4 slack, everything that is
integer or fixed point must match exactly.  Runs on the CPU."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import cases  # noqa: E402
from helpers import rel_err  # noqa: E402

FAST = 2e-4


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(HERE, "golden", "ref_kernels.npz"))


def csr_of(orc, c, random=None):
    return orc.Csr(c["start"], c["end"], c["index"], data=c["data"], weight=c["weight"], random=random)


@pytest.mark.parametrize("name,kw", cases.Z_CASES)
def test_sparse_z_vs_reference_kernels(orc, gold, name, kw):
    d = cases.dense_inputs(101)
    c = cases.make_csr(7, weighted=kw.get("weighted", False), analog=kw.get("analog", False))
    den = kw.get("denoised", False)
    rnd = d["rnd"][:len(c["index"])] if den else None
    params = orc.make_params(denoising_p=cases.DENOISE_P if den else 0.0, deltaBoost=cases.BOOST, smce=cases.SMCE)
    got = orc.sparse_z(params, csr_of(orc, c, rnd), 0, cases.BATCH, d["W"], d["Z0"].copy(), 1.0, den)
    assert rel_err(got, gold[f"z_{name}"]) < 1e-5
    np.testing.assert_array_equal(got[0], d["Z0"][0])           # the empty row is left untouched, as the reference does


def test_transposed_matrix_and_gradient_vs_reference_kernels(orc, gold):
    d = cases.dense_inputs(101)
    c = cases.make_csr(7)
    oc = csr_of(orc, c)
    tstart, cap = orc.transposed_capacity(oc, cases.WIDTH, cases.BATCH)
    params = orc.make_params()
    tend, tidx, _ = orc.sparse_transpose(params, oc, 0, cases.BATCH, tstart, cap)
    np.testing.assert_array_equal(tend, gold["t_end"])                                           # counts: bit exact
    mine = np.concatenate([np.sort(tidx[s:e]) for s, e in zip(tstart, tend)] + [np.zeros(0, np.uint32)])
    np.testing.assert_array_equal(mine, gold["t_index_sorted"])                                  # per-column sets: bit exact
    g = orc.sparse_wgrad(params, -1.0 / cases.BATCH, 0.0, tstart, tend, tidx, None, d["delta"], np.zeros((cases.WIDTH, cases.STRIDE), np.float32))
    np.testing.assert_array_equal(g, gold["wgrad"])                                              # fixed-point sums: bit exact


@pytest.mark.parametrize("act", [0, 7])
def test_activation_vs_reference_kernels(orc, gold, act):
    d = cases.dense_inputs(101)
    assert rel_err(orc.activation(act, d["z_out"].copy()), gold[f"act_{act}"]) < FAST


@pytest.mark.parametrize("ef,act,iz", cases.LOSS_CASES)
def test_loss_and_delta_vs_reference_kernels(orc, gold, ef, act, iz):
    d = cases.dense_inputs(101)
    c = cases.make_csr(11, width=cases.WIDTH)
    params = orc.make_params(deltaBoost=cases.BOOST, smce=cases.SMCE)
    unit = orc.activation(act, d["z_out"].copy())
    loss = orc.sparse_loss(params, csr_of(orc, c), ef, act, 0, cases.BATCH, unit, iz)
    delta = orc.sparse_output_delta(params, csr_of(orc, c), ef, act, 0, cases.BATCH, unit, np.zeros_like(unit), iz)
    want = float(gold[f"loss_{ef}_{act}_{int(iz)}"])
    assert abs(loss - want) <= FAST * max(abs(want), 1.0)
    assert rel_err(delta, gold[f"delta_{ef}_{act}_{int(iz)}"]) < FAST


@pytest.mark.parametrize("mode", range(7))
def test_optimizers_vs_reference_kernels(orc, gold, mode):
    d = cases.dense_inputs(101)
    hp = cases.OPT_HP
    w, v, gv = d["w"].copy(), d["v"].copy(), d["gv"].copy()
    orc.update_weights(mode, hp["alpha"], hp["lam"], hp["lam1"], hp["mu"], hp["mu1"], hp["t"], v, d["g"], gv, w)
    tol = 2e-3 if mode == orc.ADAM else FAST                  # __powf under -use_fast_math, amplified by 1/(1-beta2^t)
    assert rel_err(w, gold[f"opt_w_{mode}"]) < tol


def test_topk_vs_reference_kernel(orc, gold):
    d = cases.dense_inputs(101)
    k, v = orc.topk(d["scores"], 64)
    np.testing.assert_array_equal(k, gold["topk_key"])
    np.testing.assert_array_equal(v, gold["topk_val"])


def test_topk_vs_reference_cpu_topksort(orc):
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_topksort_golden as g
    gold = np.load(os.path.join(HERE, "golden", "ref_topksort.npz"))
    for i, (B, K, N) in enumerate(g.SHAPES):
        keys = g.keys_for(i, B, N)
        k, v = orc.topk(keys, K)
        np.testing.assert_array_equal(k, gold[f"key_{i}"])
        np.testing.assert_array_equal(v, gold[f"val_{i}"])
        if orc.ref_utils() is not None:                            # live check when oracle/_ref was built here
            rk, rv = orc.ref_topksort(keys, K)
            np.testing.assert_array_equal(k, rk)
            np.testing.assert_array_equal(v, rv)
