"""ctypes binding of the CPU oracle (oracle/liboracle.so) and of the reference-built
checkers under oracle/_ref/.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs -- never by the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

# enums (values identical to the reference's, E/NNTypes.h:65-105, E/NNEnum.h:33-45)
SGD, MOMENTUM, ADAGRAD, NESTEROV, RMSPROP, ADADELTA, ADAM = range(7)
ERR_L1, ERR_L2, ERR_CE, ERR_SMCE, ERR_DATA_SMCE, ERR_HINGE, ERR_L2HINGE = range(7)
ACT_SIGMOID, ACT_TANH, ACT_RELU, ACT_LINEAR = 0, 1, 2, 3
ACT_SOFTMAX, ACT_ELU, ACT_LRELU, ACT_SELU = 7, 10, 11, 12
DT_UINT, DT_INT, DT_LLINT, DT_ULLINT, DT_FLOAT, DT_DOUBLE, DT_UCHAR, DT_CHAR = 0, 1, 2, 3, 4, 5, 8, 9
_NP2DT = {np.dtype(np.uint32): DT_UINT, np.dtype(np.int32): DT_INT, np.dtype(np.int64): DT_LLINT,
          np.dtype(np.uint64): DT_ULLINT, np.dtype(np.float32): DT_FLOAT, np.dtype(np.float64): DT_DOUBLE,
          np.dtype(np.uint8): DT_UCHAR, np.dtype(np.int8): DT_CHAR}
MAX_VALUE = np.float32(999999999999999.0)


def build(force=False):
    """Build liboracle.so (and oracle/_ref/* when /root/reference is present)."""
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("dsstne_oracle.c", "dsstne_oracle_net.c", "dsstne_oracle.h")]
    stale = (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    if os.path.exists("/root/reference/src/amazon/dsstne/engine/kernels.cu"):
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)
    return so


class Params(C.Structure):
    _fields_ = [("bShuffleIndices", C.c_int), ("pShuffleIndex", C.c_void_p),
                ("denoising_p", C.c_float), ("denoising_q", C.c_float),
                ("deltaBoost_one", C.c_float), ("deltaBoost_zero", C.c_float),
                ("SMCE_oneTarget", C.c_float), ("SMCE_zeroTarget", C.c_float),
                ("SMCE_oneScale", C.c_float), ("SMCE_zeroScale", C.c_float)]


class CsrView(C.Structure):
    _fields_ = [("sparseStart", C.c_void_p), ("sparseEnd", C.c_void_p), ("sparseIndex", C.c_void_p),
                ("sparseData", C.c_void_p), ("dataType", C.c_int), ("dataWeight", C.c_void_p),
                ("index", C.c_void_p), ("denoisingRandom", C.c_void_p)]


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orc_sparse_loss.restype = C.c_double
        _lib.orc_regularization_error.restype = C.c_double
        _lib.orc_transposed_capacity.restype = C.c_uint32
        _lib.orc_net_create.restype = C.c_void_p
        _lib.orc_net_train_step.restype = C.c_double
        _lib.orc_net_loss.restype = C.c_double
        _lib.orc_num_threads.restype = C.c_int
    return _lib


def num_threads():
    return lib().orc_num_threads()


def set_num_threads(n):
    lib().orc_set_num_threads(C.c_int(int(n)))


def make_params(shuffle=None, denoising_p=0.0, deltaBoost=(1.0, 1.0), smce=(0.9, 0.1, 1.0, 1.0)):
    """smce = (oneTarget, zeroTarget, oneScale, zeroScale)."""
    p = Params()
    lib().orc_params_default(C.byref(p))
    if shuffle is not None:
        p.bShuffleIndices = 1
        p.pShuffleIndex = _p(shuffle)
        p._keep = shuffle
    p.denoising_p = denoising_p
    p.denoising_q = np.float32(1.0) / (np.float32(1.0) - np.float32(denoising_p))
    p.deltaBoost_one, p.deltaBoost_zero = deltaBoost
    p.SMCE_oneTarget, p.SMCE_zeroTarget, p.SMCE_oneScale, p.SMCE_zeroScale = smce
    return p


class Csr:
    """Host-side sparse dataset in DSSTNE's layout (E/NNTypes.h:213-225)."""

    def __init__(self, start, end, index, data=None, weight=None, ex_index=None, random=None):
        self.start = np.ascontiguousarray(start, dtype=np.uint64)
        self.end = np.ascontiguousarray(end, dtype=np.uint64)
        self.index = np.ascontiguousarray(index, dtype=np.uint32)
        self.data = None if data is None else np.ascontiguousarray(data)
        self.weight = None if weight is None else np.ascontiguousarray(weight, dtype=np.float32)
        self.ex_index = None if ex_index is None else np.ascontiguousarray(ex_index, dtype=np.uint32)
        self.random = None if random is None else np.ascontiguousarray(random, dtype=np.float32)

    def view(self):
        v = CsrView()
        v.sparseStart, v.sparseEnd, v.sparseIndex = _p(self.start), _p(self.end), _p(self.index)
        v.sparseData = _p(self.data)
        v.dataType = _NP2DT[self.data.dtype] if self.data is not None else DT_FLOAT
        v.dataWeight, v.index, v.denoisingRandom = _p(self.weight), _p(self.ex_index), _p(self.random)
        return v

    @property
    def examples(self):
        return len(self.ex_index) if self.ex_index is not None else len(self.start)

    @property
    def unique_examples(self):
        return len(self.start)


def _f32(a):
    assert a.dtype == np.float32 and a.flags.c_contiguous
    return _p(a)


def clear_unit(unit, bias):
    batch, stride = unit.shape
    lib().orc_clear_unit(_f32(unit), _f32(bias), C.c_uint32(stride), C.c_uint32(batch))


def sparse_z(params, csr, position, batch, W, Z, beta=1.0, denoised=False):
    stride = W.shape[1]
    v = csr.view()
    lib().orc_sparse_z(C.byref(params), C.byref(v), C.c_uint32(position), C.c_uint32(batch), C.c_uint32(stride),
                       _f32(W), _f32(Z), C.c_float(beta), C.c_int(int(denoised)))
    return Z


def transposed_capacity(csr, N, batch):
    tstart = np.zeros(N, dtype=np.uint32)
    v = csr.view()
    cap = lib().orc_transposed_capacity(C.byref(v), C.c_uint32(csr.examples), C.c_uint32(csr.unique_examples),
                                        C.c_uint32(N), C.c_uint32(batch), _p(tstart))
    return tstart, int(cap)


def sparse_transpose(params, csr, position, batch, tstart, cap, denoised=False, with_data=None):
    if with_data is None:
        with_data = (csr.data is not None) or (csr.weight is not None)
    tend = tstart.copy()
    tindex = np.zeros(max(cap, 1), dtype=np.uint32)
    tdata = np.zeros(max(cap, 1), dtype=np.float32) if with_data else None
    v = csr.view()
    lib().orc_sparse_transpose(C.byref(params), C.byref(v), C.c_uint32(position), C.c_uint32(batch),
                               C.c_int(int(denoised)), _p(tend), _p(tindex), _p(tdata))
    return tend, tindex, tdata


def sparse_wgrad(params, alpha, beta, tstart, tend, tindex, tdata, delta, dW):
    m, n = dW.shape
    lib().orc_sparse_wgrad(C.byref(params), C.c_float(alpha), C.c_float(beta), C.c_uint32(m), C.c_uint32(n),
                           _p(tstart), _p(tend), _p(tindex), _p(tdata), _f32(delta), _f32(dW))
    return dW


def activation(act, data, slope=0.0, alpha=0.0, lam=0.0):
    batch, stride = data.shape
    lib().orc_activation(C.c_int(act), _f32(data), C.c_uint32(batch), C.c_uint32(stride),
                         C.c_float(slope), C.c_float(alpha), C.c_float(lam))
    return data


def sparse_loss(params, csr, ef, act, position, batch, unit, ignore_zero=False):
    stride = unit.shape[1]
    v = csr.view()
    return lib().orc_sparse_loss(C.byref(params), C.byref(v), C.c_int(ef), C.c_int(act), C.c_uint32(position),
                                 C.c_uint32(batch), C.c_uint32(stride), _f32(unit), C.c_int(int(ignore_zero)))


def sparse_output_delta(params, csr, ef, act, position, batch, unit, delta, ignore_zero=False,
                        slope=0.0, alpha=0.0, lam=0.0):
    stride = unit.shape[1]
    v = csr.view()
    lib().orc_sparse_output_delta(C.byref(params), C.byref(v), C.c_int(ef), C.c_int(act), C.c_uint32(position),
                                  C.c_uint32(batch), C.c_uint32(stride), _f32(unit), _f32(delta),
                                  C.c_int(int(ignore_zero)), C.c_float(slope), C.c_float(alpha), C.c_float(lam))
    return delta


def sparseness_penalty(unit, delta, p, beta):
    batch, stride = unit.shape
    lib().orc_sparseness_penalty(C.c_uint32(batch), C.c_uint32(stride), _f32(unit), _f32(delta), C.c_float(p), C.c_float(beta))
    return delta


def hadamard(act, unit, delta, scale=1.0, slope=0.0, alpha=0.0, lam=0.0):
    lib().orc_hadamard(C.c_int(act), C.c_uint64(unit.size), C.c_float(scale), _f32(unit), _f32(delta),
                       C.c_float(slope), C.c_float(alpha), C.c_float(lam))
    return delta


def dropout(act, unit, random, p, elu_alpha=1.0, selu_lambda=1.050701):
    batch, stride = unit.shape
    lib().orc_dropout(C.c_int(act), _f32(unit), _f32(random), C.c_uint32(batch), C.c_uint32(stride), C.c_float(p),
                      C.c_float(elu_alpha), C.c_float(selu_lambda))
    return unit


def gemm_fwd(A, W, C_, beta=1.0):
    B, k = A.shape
    n = W.shape[1]
    lib().orc_gemm_fwd(C.c_uint32(B), C.c_uint32(k), C.c_uint32(n), _f32(A), _f32(W), C.c_float(beta), _f32(C_))
    return C_


def gemm_dw(A, D, G, alpha, beta=0.0):
    B, k = A.shape
    n = D.shape[1]
    lib().orc_gemm_dw(C.c_uint32(B), C.c_uint32(k), C.c_uint32(n), C.c_float(alpha), _f32(A), _f32(D), C.c_float(beta), _f32(G))
    return G


def gemm_dx(D, W, Dp, beta=0.0):
    B, n = D.shape
    k = W.shape[0]
    lib().orc_gemm_dx(C.c_uint32(B), C.c_uint32(k), C.c_uint32(n), _f32(D), _f32(W), C.c_float(beta), _f32(Dp))
    return Dp


def update_weights(mode, alpha, lam, lam1, mu, mu1, t, v, g, gv, w):
    lib().orc_update_weights(C.c_int(mode), C.c_float(alpha), C.c_float(lam), C.c_float(lam1), C.c_float(mu),
                             C.c_float(mu1), C.c_float(t), C.c_uint64(w.size), _p(v), _f32(g), _p(gv), _f32(w))


def update_biases(mode, alpha, mu, mu1, t, delta, v, gv, bias):
    batch, width = delta.shape
    lib().orc_update_biases(C.c_int(mode), C.c_float(alpha), C.c_float(mu), C.c_float(mu1), C.c_float(t),
                            C.c_uint32(batch), C.c_uint32(width), _f32(delta), _p(v), _p(gv), _f32(bias))


def regularization_error(lam, lam1, w):
    return lib().orc_regularization_error(C.c_float(lam), C.c_float(lam1), _f32(w), C.c_uint64(w.size))


def topk(key, k, value=None, filt=None):
    """filt = (start u64[B], end u64[B], index u32[nnz]) or None."""
    batch, width = key.shape
    out_key = np.empty((batch, k), dtype=np.float32)
    out_val = np.empty((batch, k), dtype=np.uint32)
    fs = fe = fi = None
    if filt is not None:
        fs, fe, fi = (np.ascontiguousarray(filt[0], dtype=np.uint64), np.ascontiguousarray(filt[1], dtype=np.uint64),
                      np.ascontiguousarray(filt[2], dtype=np.uint32))
    lib().orc_topk(_f32(key), _p(value), C.c_uint32(batch), C.c_uint32(width), C.c_uint32(k),
                   _p(fs), _p(fe), _p(fi), _p(out_key), _p(out_val))
    return out_key, out_val


def shard_range(N, rank, nranks):
    a, b = C.c_uint32(), C.c_uint32()
    lib().orc_shard_range(C.c_uint32(N), C.c_uint32(rank), C.c_uint32(nranks), C.byref(a), C.byref(b))
    return a.value, b.value


def weight_outgoing_larger(in_stride, out_stride):
    return bool(lib().orc_weight_outgoing_larger(C.c_uint32(in_stride), C.c_uint32(out_stride)))


class _NetStruct(C.Structure):
    _MAXW = 8
    _fields_ = [("nWeights", C.c_int), ("size", C.c_uint32 * 9), ("activation", C.c_int * 9),
                ("sparsePenalty", C.c_int * 9), ("errorFunction", C.c_int), ("trainingMode", C.c_int),
                ("denoising", C.c_int), ("sparsenessPenalty_p", C.c_float), ("sparsenessPenalty_beta", C.c_float),
                ("params", Params), ("maxBatch", C.c_uint32), ("batches", C.c_uint64),
                ("W", C.POINTER(C.c_float) * 8), ("b", C.POINTER(C.c_float) * 8), ("dW", C.POINTER(C.c_float) * 8),
                ("vW", C.POINTER(C.c_float) * 8), ("gvW", C.POINTER(C.c_float) * 8),
                ("vb", C.POINTER(C.c_float) * 8), ("gvb", C.POINTER(C.c_float) * 8),
                ("unit", C.POINTER(C.c_float) * 9), ("delta", C.POINTER(C.c_float) * 9),
                ("tStart", C.POINTER(C.c_uint32)), ("tEnd", C.POINTER(C.c_uint32)),
                ("tIndex", C.POINTER(C.c_uint32)), ("tData", C.POINTER(C.c_float)), ("tCapacity", C.c_uint32),
                ("pDropout", C.c_float * 9), ("dropoutRandom", C.POINTER(C.c_float) * 9)]


class Network:
    """Sparse-in / sparse-out FC network on the CPU oracle (orc_net_*)."""

    def __init__(self, sizes, activations=None, error=ERR_SMCE, mode=SGD, max_batch=256):
        L = len(sizes) - 1
        acts = [ACT_SIGMOID] * (L + 1) if activations is None else [ACT_SIGMOID] + list(activations)
        sz = (C.c_uint32 * (L + 1))(*sizes)
        ac = (C.c_int * (L + 1))(*acts)
        self._h = lib().orc_net_create(C.c_int(L), sz, ac, C.c_int(error), C.c_int(mode), C.c_uint32(max_batch))
        assert self._h
        self.s = C.cast(self._h, C.POINTER(_NetStruct)).contents
        self.sizes = list(sizes)
        self.L = L
        self.max_batch = max_batch

    def close(self):
        if self._h:
            lib().orc_net_destroy(C.c_void_p(self._h))
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _arr(self, ptr, shape):
        n = int(np.prod(shape))
        return np.ctypeslib.as_array(ptr, shape=(n,)).reshape(shape)

    def W(self, i):
        return self._arr(self.s.W[i], (self.sizes[i], self.sizes[i + 1]))

    def b(self, i):
        return self._arr(self.s.b[i], (self.sizes[i + 1],))

    def dW(self, i):
        return self._arr(self.s.dW[i], (self.sizes[i], self.sizes[i + 1]))

    def unit(self, l, batch):
        return self._arr(self.s.unit[l], (self.max_batch, self.sizes[l]))[:batch]

    def delta(self, l, batch):
        return self._arr(self.s.delta[l], (self.max_batch, self.sizes[l]))[:batch]

    def set_dropout(self, l, p, random):
        """hidden layer l drops with probability p using the uniform randoms `random` [max_batch][size[l]] (kept alive here)."""
        self._drop = getattr(self, "_drop", {})
        self._drop[l] = np.ascontiguousarray(random, dtype=np.float32)
        self.s.pDropout[l] = p
        self.s.dropoutRandom[l] = self._drop[l].ctypes.data_as(C.POINTER(C.c_float))

    def set_input(self, csr, batch):
        self._in_view = csr.view()
        self._in = csr
        lib().orc_net_set_input(C.c_void_p(self._h), C.byref(self._in_view), C.c_uint32(csr.examples),
                                C.c_uint32(csr.unique_examples), C.c_uint32(batch))

    def forward(self, csr_in, position, batch, training=False):
        vi = csr_in.view()
        lib().orc_net_forward(C.c_void_p(self._h), C.byref(vi), C.c_uint32(position), C.c_uint32(batch), C.c_int(int(training)))

    def loss(self, csr_in, csr_out, position, batch):
        vi, vo = csr_in.view(), csr_out.view()
        return lib().orc_net_loss(C.c_void_p(self._h), C.byref(vi), C.byref(vo), C.c_uint32(position), C.c_uint32(batch))

    def backward(self, csr_in, csr_out, position, batch):
        vi, vo = csr_in.view(), csr_out.view()
        lib().orc_net_backward(C.c_void_p(self._h), C.byref(vi), C.byref(vo), C.c_uint32(position), C.c_uint32(batch))

    def train_step(self, csr_in, csr_out, position, batch, alpha, lam=0.0, lam1=0.0, mu=0.0, mu1=0.0):
        vi, vo = csr_in.view(), csr_out.view()
        reg = C.c_double()
        e = lib().orc_net_train_step(C.c_void_p(self._h), C.byref(vi), C.byref(vo), C.c_uint32(position),
                                     C.c_uint32(batch), C.c_float(alpha), C.c_float(lam), C.c_float(lam1),
                                     C.c_float(mu), C.c_float(mu1), C.byref(reg))
        return e, reg.value


# ---- reference-built checkers (present when oracle/_ref was built from /root/reference) ----
def ref_utils():
    """The reference's own CPU top-K comparator (U/Utils.cpp:213-243), or None."""
    path = os.path.join(_HERE, "_ref", "libdsstne_refutils.so")
    if not os.path.exists(path):
        return None
    return C.CDLL(path)


def ref_topksort(keys, k):
    """Run the reference topKsort<float,uint32> row by row. keys: [B][N] float32."""
    l = ref_utils()
    assert l is not None
    B, N = keys.shape
    ok = np.empty((B, k), dtype=np.float32)
    ov = np.empty((B, k), dtype=np.uint32)
    for b in range(B):
        row = np.ascontiguousarray(keys[b])
        l.ref_topKsort_f32_u32(_p(row), None, C.c_int(N), _p(ok[b]), _p(ov[b]), C.c_int(k))
    return ok, ov


def ref_kernels():
    """The reference's own CUDA kernels built for sm_100 (needs a GPU), or None."""
    path = os.path.join(_HERE, "_ref", "libdsstne_refkernels.so")
    if not os.path.exists(path):
        return None
    l = C.CDLL(path)
    l.ref_sparse_loss.restype = C.c_float
    l.ref_regularization_error.restype = C.c_float
    return l
