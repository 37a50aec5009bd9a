"""dsstne_b200 -- Python-side loader for libdsstne_b200.so (the C ABI of include/dsstne_b200.h).

The product is the shared library (hand-written sm_100a kernels behind a C ABI that replaces
DSSTNE's E/kernels.h) and the C++ host mirror of NNDataSet/NNLayer/NNWeight/NNNetwork in
`engine/`.  This module only binds it with ctypes for tests and bench.py; torch is used for
device memory and streams (plumbing).  There is NO CPU fallback: loading fails loudly when the
extension is missing, and every compute call needs a B200.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdsstne_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "dsstne_b200.h")

# enums (include/dsstne_b200.h)
SGD, MOMENTUM, ADAGRAD, NESTEROV, RMSPROP, ADADELTA, ADAM = range(7)
ERR_L1, ERR_L2, ERR_CE, ERR_SMCE, ERR_DATA_SMCE, ERR_HINGE, ERR_L2HINGE = range(7)
ACT_SIGMOID, ACT_TANH, ACT_RELU, ACT_LINEAR = 0, 1, 2, 3
ACT_SOFTMAX, ACT_ELU, ACT_LRELU, ACT_SELU = 7, 10, 11, 12
DT_UINT, DT_INT, DT_LLINT, DT_ULLINT, DT_FLOAT, DT_DOUBLE, DT_UCHAR, DT_CHAR = 0, 1, 2, 3, 4, 5, 8, 9
GEMM_FP32, GEMM_TF32, GEMM_TF32X3 = 0, 1, 2


class DsbError(RuntimeError):
    pass


def build(force=False):
    """Compile libdsstne_b200.so for sm_100a (nvcc cross-compiles without a GPU)."""
    if force:
        subprocess.check_call(["make", "-C", _HERE, "clean"], stdout=subprocess.DEVNULL)
    subprocess.check_call(["make", "-C", _HERE, "-j8"], stdout=subprocess.DEVNULL)
    return LIB_PATH


class Params(C.Structure):
    _fields_ = [("bShuffleIndices", C.c_int32), ("pShuffleIndex", C.c_void_p),
                ("denoising_p", C.c_float), ("denoising_q", C.c_float),
                ("deltaBoost_one", C.c_float), ("deltaBoost_zero", C.c_float),
                ("SMCE_oneTarget", C.c_float), ("SMCE_zeroTarget", C.c_float),
                ("SMCE_oneScale", C.c_float), ("SMCE_zeroScale", C.c_float)]


class Sparse(C.Structure):
    _fields_ = [("sparseStart", C.c_void_p), ("sparseEnd", C.c_void_p), ("sparseIndex", C.c_void_p),
                ("sparseData", C.c_void_p), ("dataType", C.c_int32), ("dataWeight", C.c_void_p),
                ("index", C.c_void_p), ("denoisingRandom", C.c_void_p)]


_lib = None


def lib():
    """The loaded C-ABI library; raises if the extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise DsbError(f"{LIB_PATH} is missing: run __graft_entry__.build() (there is no CPU fallback)")
        _lib = C.CDLL(LIB_PATH)        # RTLD_LOCAL: the engine exports C++ names (getGpu, GpuContext) the reference also uses
        _lib.dsb200_last_error.restype = C.c_char_p
        _lib.dsb200_launch_count.restype = C.c_uint64
    return _lib


def _ptr(t):
    """Device pointer of a torch tensor (or None)."""
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


class Context:
    """Owns a dsb200_ctx bound to the current torch CUDA stream."""

    def __init__(self, device=0):
        import torch
        self._torch = torch
        self.device = device
        self.h = C.c_void_p()
        rc = lib().dsb200_ctx_create(C.byref(self.h), C.c_int(device))
        if rc:
            raise DsbError(f"dsb200_ctx_create failed ({rc}): no sm_100 GPU -- there is no CPU fallback")
        self.use_current_stream()
        self.params = Params()
        lib().dsb200_params_default(C.byref(self.params))
        self._keep = []

    def close(self):
        if self.h:
            lib().dsb200_ctx_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc):
        if rc:
            raise DsbError(f"dsb200 error {rc}: {lib().dsb200_last_error(self.h).decode()}")

    def use_current_stream(self):
        s = self._torch.cuda.current_stream(self.device).cuda_stream
        self.check(lib().dsb200_ctx_set_stream(self.h, C.c_void_p(s)))

    def set_params(self, shuffle=None, denoising_p=0.0, deltaBoost=(1.0, 1.0), smce=(0.9, 0.1, 1.0, 1.0)):
        import numpy as np
        p = self.params
        p.bShuffleIndices = 0 if shuffle is None else 1
        p.pShuffleIndex = None if shuffle is None else shuffle.data_ptr()
        self._keep = [shuffle]
        p.denoising_p = denoising_p
        p.denoising_q = float(np.float32(1.0) / (np.float32(1.0) - np.float32(denoising_p)))
        p.deltaBoost_one, p.deltaBoost_zero = deltaBoost
        p.SMCE_oneTarget, p.SMCE_zeroTarget, p.SMCE_oneScale, p.SMCE_zeroScale = smce
        self.check(lib().dsb200_ctx_set_params(self.h, C.byref(p)))

    def set_option(self, name, value):
        self.check(lib().dsb200_ctx_set_option(self.h, name.encode(), C.c_int(int(value))))

    def sync(self):
        self.check(lib().dsb200_ctx_sync(self.h))

    # ---- kernel families -------------------------------------------------------------
    def clear_unit(self, unit, bias):
        b, s = unit.shape
        self.check(lib().dsb200_clear_unit(self.h, _ptr(unit), _ptr(bias), C.c_uint32(s), C.c_uint32(b)))

    def add_bias(self, unit, bias):
        b, s = unit.shape
        self.check(lib().dsb200_add_bias(self.h, _ptr(unit), _ptr(bias), C.c_uint32(s), C.c_uint32(b)))

    def sparse_z(self, ds, position, batch, W, Z, beta=1.0, denoised=False):
        v = ds.view()
        self.check(lib().dsb200_sparse_z(self.h, C.byref(v), C.c_uint32(position), C.c_uint32(batch),
                                         C.c_uint32(W.shape[1]), _ptr(W), _ptr(Z), C.c_float(beta),
                                         C.c_int(int(denoised))))

    def sparse_z_bias_act(self, ds, position, batch, W, bias, act, unit, denoised=False):
        v = ds.view()
        self.check(lib().dsb200_sparse_z_bias_act(self.h, C.byref(v), C.c_uint32(position), C.c_uint32(batch),
                                                  C.c_uint32(W.shape[1]), _ptr(W), _ptr(bias), C.c_int(act),
                                                  _ptr(unit), C.c_int(int(denoised))))

    def sparse_transpose(self, ds, position, batch, N, tstart, tend, tindex, tdata=None, denoised=False):
        v = ds.view()
        self.check(lib().dsb200_sparse_transpose(self.h, C.byref(v), C.c_uint32(position), C.c_uint32(batch),
                                                 C.c_int(int(denoised)), C.c_uint32(N), _ptr(tstart), _ptr(tend),
                                                 _ptr(tindex), _ptr(tdata)))

    def transposed_capacity(self, ds, rows, N, count, tstart, total=None):
        v = ds.view()
        self.check(lib().dsb200_transposed_capacity(self.h, C.byref(v), C.c_uint32(rows), C.c_uint32(N), _ptr(count), _ptr(tstart), _ptr(total)))

    def sparse_wgrad(self, alpha, beta, tstart, tend, tindex, tdata, delta, dW):
        m, n = dW.shape
        self.check(lib().dsb200_sparse_wgrad(self.h, C.c_float(alpha), C.c_float(beta), C.c_uint32(m), C.c_uint32(n),
                                             _ptr(tstart), _ptr(tend), _ptr(tindex), _ptr(tdata), _ptr(delta), _ptr(dW)))

    def sparse_wgrad_update(self, mode, galpha, tstart, tend, tindex, tdata, delta, alpha, lam, lam1, mu, mu1, t, v, gv, w):
        m, n = w.shape
        self.check(lib().dsb200_sparse_wgrad_update(self.h, C.c_int(mode), C.c_float(galpha), C.c_uint32(m), C.c_uint32(n),
                                                    _ptr(tstart), _ptr(tend), _ptr(tindex), _ptr(tdata), _ptr(delta),
                                                    C.c_float(alpha), C.c_float(lam), C.c_float(lam1), C.c_float(mu),
                                                    C.c_float(mu1), C.c_float(t), _ptr(v), _ptr(gv), _ptr(w)))

    def activation(self, act, data, slope=0.0, alpha=0.0, lam=0.0):
        b, s = data.shape
        self.check(lib().dsb200_activation(self.h, C.c_int(act), _ptr(data), C.c_uint32(b), C.c_uint32(s),
                                           C.c_float(slope), C.c_float(alpha), C.c_float(lam)))

    def sparse_loss(self, ds, ef, act, position, batch, unit, ignore_zero=False):
        v = ds.view()
        out = C.c_float()
        self.check(lib().dsb200_sparse_loss(self.h, C.byref(v), C.c_int(ef), C.c_int(act), C.c_uint32(position),
                                            C.c_uint32(batch), C.c_uint32(unit.shape[1]), _ptr(unit),
                                            C.c_int(int(ignore_zero)), C.byref(out)))
        return out.value

    def sparse_output_delta(self, ds, ef, act, position, batch, unit, delta, ignore_zero=False, slope=0.0, alpha=0.0, lam=0.0):
        v = ds.view()
        self.check(lib().dsb200_sparse_output_delta(self.h, C.byref(v), C.c_int(ef), C.c_int(act), C.c_uint32(position),
                                                    C.c_uint32(batch), C.c_uint32(unit.shape[1]), _ptr(unit), _ptr(delta),
                                                    C.c_int(int(ignore_zero)), C.c_float(slope), C.c_float(alpha), C.c_float(lam)))

    def output_pass(self, ds, ef, act, position, batch, z, unit_out, delta, acc=None):
        v = ds.view()
        self.check(lib().dsb200_output_pass(self.h, C.byref(v), C.c_int(ef), C.c_int(act), C.c_uint32(position),
                                            C.c_uint32(batch), C.c_uint32(z.shape[1]), _ptr(z), _ptr(unit_out),
                                            _ptr(delta), _ptr(acc)))

    def gemm_fwd_output_pass(self, ds, ef, act, position, A, W, bias, unit_out, delta, acc=None, col_partials=None):
        """Forward GEMM of a sigmoid output layer with loss + delta in its epilogue (Z is never written).  Returns the number of
        rows of column-sum partials written into col_partials ([4 * ceil(B / 128)][n], optional)."""
        v = ds.view()
        B, k = A.shape
        n_part = C.c_uint32(0)
        self.check(lib().dsb200_gemm_fwd_output_pass(self.h, C.byref(v), C.c_int(ef), C.c_int(act), C.c_uint32(position), C.c_uint32(B),
                                                     C.c_uint32(k), C.c_uint32(W.shape[1]), _ptr(A), _ptr(W), _ptr(bias), _ptr(unit_out),
                                                     _ptr(delta), _ptr(acc), _ptr(col_partials), C.byref(n_part)))
        return n_part.value

    def update_biases_partials(self, mode, alpha, mu, mu1, t, batch, partials, n_partials, v, gv, bias):
        self.check(lib().dsb200_update_biases_partials(self.h, C.c_int(mode), C.c_float(alpha), C.c_float(mu), C.c_float(mu1), C.c_float(t),
                                                       C.c_uint32(batch), C.c_uint32(bias.numel()), _ptr(partials), C.c_uint32(n_partials),
                                                       _ptr(v), _ptr(gv), _ptr(bias)))

    def sparseness_penalty(self, unit, delta, p, beta):
        b, s = unit.shape
        self.check(lib().dsb200_sparseness_penalty(self.h, C.c_uint32(b), C.c_uint32(s), _ptr(unit), _ptr(delta),
                                                   C.c_float(p), C.c_float(beta)))

    def hadamard(self, act, unit, delta, scale=1.0, slope=0.0, alpha=0.0, lam=0.0):
        self.check(lib().dsb200_hadamard(self.h, C.c_int(act), C.c_uint64(unit.numel()), C.c_float(scale), _ptr(unit),
                                         _ptr(delta), C.c_float(slope), C.c_float(alpha), C.c_float(lam)))

    def dropout(self, act, unit, p, seed, stream, full_stride=None, col_offset=0, elu_alpha=1.0, selu_lambda=1.050701):
        b, s = unit.shape
        self.check(lib().dsb200_dropout(self.h, C.c_int(act), _ptr(unit), C.c_uint32(b), C.c_uint32(s), C.c_uint32(full_stride or s),
                                        C.c_uint32(col_offset), C.c_float(p), C.c_float(elu_alpha), C.c_float(selu_lambda),
                                        C.c_uint64(seed), C.c_uint64(stream)))

    def gemm_fwd(self, A, W, Cm, beta=1.0):
        B, k = A.shape
        self.check(lib().dsb200_gemm_fwd(self.h, C.c_uint32(B), C.c_uint32(k), C.c_uint32(W.shape[1]), _ptr(A), _ptr(W),
                                         C.c_float(beta), _ptr(Cm)))

    def gemm_fwd_bias_act(self, A, W, bias, act, Cm, slope=0.0, alpha=0.0, lam=0.0):
        B, k = A.shape
        self.check(lib().dsb200_gemm_fwd_bias_act(self.h, C.c_uint32(B), C.c_uint32(k), C.c_uint32(W.shape[1]), _ptr(A), _ptr(W),
                                                  _ptr(bias), C.c_int(act), _ptr(Cm), C.c_float(slope), C.c_float(alpha), C.c_float(lam)))

    def gemm_dx_hadamard(self, D, W, act, unit, Dp, scale=1.0, slope=0.0, alpha=0.0, lam=0.0):
        B, n = D.shape
        self.check(lib().dsb200_gemm_dx_hadamard(self.h, C.c_uint32(B), C.c_uint32(W.shape[0]), C.c_uint32(n), _ptr(D), _ptr(W), C.c_int(act),
                                                 C.c_float(scale), _ptr(unit), _ptr(Dp), C.c_float(slope), C.c_float(alpha), C.c_float(lam)))

    def gemm_dw(self, A, D, G, alpha, beta=0.0):
        B, k = A.shape
        self.check(lib().dsb200_gemm_dw(self.h, C.c_uint32(B), C.c_uint32(k), C.c_uint32(D.shape[1]), C.c_float(alpha),
                                        _ptr(A), _ptr(D), C.c_float(beta), _ptr(G)))

    def gemm_dx(self, D, W, Dp, beta=0.0):
        B, n = D.shape
        self.check(lib().dsb200_gemm_dx(self.h, C.c_uint32(B), C.c_uint32(W.shape[0]), C.c_uint32(n), _ptr(D), _ptr(W),
                                        C.c_float(beta), _ptr(Dp)))

    def update_weights(self, mode, alpha, lam, lam1, mu, mu1, t, v, g, gv, w):
        self.check(lib().dsb200_update_weights(self.h, C.c_int(mode), C.c_float(alpha), C.c_float(lam), C.c_float(lam1),
                                               C.c_float(mu), C.c_float(mu1), C.c_float(t), C.c_uint64(w.numel()),
                                               _ptr(v), _ptr(g), _ptr(gv), _ptr(w)))

    def update_biases(self, mode, alpha, mu, mu1, t, delta, v, gv, bias):
        b, width = delta.shape
        self.check(lib().dsb200_update_biases(self.h, C.c_int(mode), C.c_float(alpha), C.c_float(mu), C.c_float(mu1),
                                              C.c_float(t), C.c_uint32(b), C.c_uint32(width), _ptr(delta), _ptr(v),
                                              _ptr(gv), _ptr(bias)))

    def dense_update(self, mode, galpha, X, D, alpha, lam, lam1, mu, mu1, t, v, gv, w, bv, bgv, bias):
        """small dense layer: X^T * D weight gradient + optimizer rule + bias update in one launch (csrc/dense_small.cu)"""
        B, k = X.shape
        self.check(lib().dsb200_dense_update(self.h, C.c_int(mode), C.c_uint32(B), C.c_uint32(k), C.c_uint32(D.shape[1]), C.c_float(galpha),
                                             _ptr(X), _ptr(D), C.c_float(alpha), C.c_float(lam), C.c_float(lam1), C.c_float(mu), C.c_float(mu1),
                                             C.c_float(t), _ptr(v), _ptr(gv), _ptr(w), _ptr(bv), _ptr(bgv), _ptr(bias)))

    def regularization_error(self, lam, lam1, w):
        out = C.c_float()
        self.check(lib().dsb200_regularization_error(self.h, C.c_float(lam), C.c_float(lam1), _ptr(w),
                                                     C.c_uint64(w.numel()), C.byref(out)))
        return out.value

    def topk(self, scores, k, out_key, out_val, filt=None):
        b, width = scores.shape
        fs, fe, fi = (None, None, None) if filt is None else filt
        self.check(lib().dsb200_topk(self.h, _ptr(scores), C.c_uint32(b), C.c_uint32(width), C.c_uint32(k),
                                     _ptr(fs), _ptr(fe), _ptr(fi), _ptr(out_key), _ptr(out_val)))

    def topk_kv(self, key, value, k, out_key, out_val):
        b, width = key.shape
        self.check(lib().dsb200_topk_kv(self.h, _ptr(key), _ptr(value), C.c_uint32(b), C.c_uint32(width), C.c_uint32(k),
                                        _ptr(out_key), _ptr(out_val)))

    def topk_offset(self, value, offset):
        """value[i] += offset in place (local column ids -> global ids before the cross-rank merge)."""
        self.check(lib().dsb200_topk_offset(self.h, _ptr(value), C.c_uint64(value.numel()), C.c_uint32(offset)))


class DeviceCsr:
    """A sparse dataset resident in HBM, in DSSTNE's layout (E/NNTypes.h:213-236)."""

    _NP2DT = None

    def __init__(self, start, end, index, data=None, weight=None, ex_index=None, random=None, device="cuda:0"):
        import numpy as np
        import torch
        if DeviceCsr._NP2DT is None:
            DeviceCsr._NP2DT = {np.dtype(np.uint32): DT_UINT, np.dtype(np.int32): DT_INT, np.dtype(np.int64): DT_LLINT,
                                np.dtype(np.uint64): DT_ULLINT, np.dtype(np.float32): DT_FLOAT,
                                np.dtype(np.float64): DT_DOUBLE, np.dtype(np.uint8): DT_UCHAR, np.dtype(np.int8): DT_CHAR}

        def up(a, dt):
            if a is None:
                return None
            a = np.ascontiguousarray(a, dtype=dt)
            # torch has no uint32/uint64 arithmetic, but we only need the bytes on the device;
            # pad to >= 16 bytes so empty arrays still have a valid device address
            raw = a.view(np.uint8).reshape(-1)
            if raw.size < 16:
                raw = np.concatenate([raw, np.zeros(16 - raw.size, dtype=np.uint8)])
            return torch.from_numpy(raw.copy()).to(device)

        self.n_rows = len(start)
        self.nnz = len(index)
        self.start = up(start, np.uint64)
        self.end = up(end, np.uint64)
        self.index = up(index, np.uint32)
        self.data_type = DT_FLOAT
        self.data = None
        if data is not None:
            d = np.ascontiguousarray(data)
            self.data_type = DeviceCsr._NP2DT[d.dtype]
            self.data = up(d, d.dtype)
        self.weight = up(weight, np.float32)
        self.ex_index = up(ex_index, np.uint32)
        self.random = up(random, np.float32)

    def view(self):
        v = Sparse()
        v.sparseStart = self.start.data_ptr()
        v.sparseEnd = self.end.data_ptr()
        v.sparseIndex = self.index.data_ptr()
        v.sparseData = None if self.data is None else self.data.data_ptr()
        v.dataType = self.data_type
        v.dataWeight = None if self.weight is None else self.weight.data_ptr()
        v.index = None if self.ex_index is None else self.ex_index.data_ptr()
        v.denoisingRandom = None if self.random is None else self.random.data_ptr()
        return v
