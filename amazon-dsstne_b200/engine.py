"""ctypes binding of the network-level C ABI (include/dsstne_b200_engine.h): the C++ host mirror of
NNNetwork / NNDataSet that lives in engine/.  HOST numpy arrays in and out; the engine owns the
device copies (as NNDataSet does in the reference).  Used by bench.py, smoke() and the tests."""
import ctypes as C

import numpy as np

from . import lib, DsbError

_DT = {np.dtype(np.uint32): 0, np.dtype(np.int32): 1, np.dtype(np.float32): 4, np.dtype(np.float64): 5,
       np.dtype(np.uint8): 8, np.dtype(np.int8): 9}


def _check(rc):
    if rc:
        lib().dsb200_engine_last_error.restype = C.c_char_p
        raise DsbError(f"dsb200 engine error {rc}: {lib().dsb200_engine_last_error().decode()}")


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def unique_id():
    """128-byte NCCL unique id (rank 0 creates it; the launcher broadcasts it)."""
    buf = (C.c_char * 128)()
    rc = lib().dsb200_comm_unique_id(buf)
    if rc:
        raise DsbError(f"dsb200_comm_unique_id failed ({rc})")
    return bytes(buf)


def startup(rank=0, nranks=1, device=0, nccl_id=None, seed=12134, stream=None):
    """getGpu().Startup + SetRandomSeed (U/Train.cpp:118-119)."""
    idbuf = None
    if nccl_id is not None:
        idbuf = C.create_string_buffer(bytes(nccl_id), 128)
    _check(lib().dsb200_engine_startup(C.c_int(rank), C.c_int(nranks), C.c_int(device), idbuf, C.c_uint64(seed)))
    if stream is not None:
        set_stream(stream)


def set_stream(stream):
    _check(lib().dsb200_engine_set_stream(C.c_void_p(stream)))


def use_torch_stream(device=0):
    import torch
    set_stream(torch.cuda.current_stream(device).cuda_stream)


def shutdown():
    _check(lib().dsb200_engine_shutdown())


def set_option(name, value):
    _check(lib().dsb200_engine_set_option(name.encode(), C.c_int(int(value))))


def profile_report():
    """{family: (calls, total_ms)} accumulated since the last report (option "profile" must be 1)."""
    buf = C.create_string_buffer(1 << 16)
    _check(lib().dsb200_engine_profile_report(buf, C.c_size_t(len(buf))))
    out = {}
    for line in buf.value.decode().splitlines():
        name, calls, ms = line.split()
        out[name] = (int(calls), float(ms))
    return out


def sync():
    _check(lib().dsb200_engine_sync())


def describe_json(json_text, datasets=None):
    """HOST ONLY: what the JSON (LDL) loader understands of a network description, one line per network / layer / weight.
    `datasets` = {name: width} for auto-sized layers.  Raises EngineError exactly where LoadNeuralNetworkJSON would fail."""
    datasets = datasets or {}
    names = (C.c_char_p * len(datasets))(*[n.encode() for n in datasets])
    widths = (C.c_uint32 * len(datasets))(*[int(w) for w in datasets.values()])
    buf = C.create_string_buffer(1 << 16)
    _check(lib().dsb200_describe_network_json(json_text.encode(), names, widths, C.c_int(len(datasets)), buf, C.c_size_t(len(buf))))
    return buf.value.decode()


class Dataset:
    """NNDataSet<T> (sparse): Boolean when data is None."""

    def __init__(self, name, start, end, index, width, data=None, weight=None, ex_index=None, ignore_zero=False):
        self.start = np.ascontiguousarray(start, dtype=np.uint64)
        self.end = np.ascontiguousarray(end, dtype=np.uint64)
        self.index = np.ascontiguousarray(index, dtype=np.uint32)
        self.data = None if data is None else np.ascontiguousarray(data)
        self.weight = None if weight is None else np.ascontiguousarray(weight, dtype=np.float32)
        self.ex_index = None if ex_index is None else np.ascontiguousarray(ex_index, dtype=np.uint32)
        unique = len(self.start)
        examples = len(self.ex_index) if self.ex_index is not None else unique
        dt = 4 if self.data is None else _DT[self.data.dtype]
        self.h = C.c_void_p()
        self.name = name
        self.width = width
        self.examples = examples
        _check(lib().dsb200_dataset_create_sparse(C.byref(self.h), name.encode(), C.c_int(dt), C.c_uint32(examples),
                                                  C.c_uint32(unique), C.c_uint32(width), C.c_uint32(1), C.c_uint32(1),
                                                  _p(self.start), _p(self.end), _p(self.index), _p(self.data),
                                                  _p(self.weight), _p(self.ex_index), C.c_int(int(ignore_zero))))

    @classmethod
    def from_host_csr(cls, name, h, **kw):
        return cls(name, h.start, h.end, h.index, h.width, data=h.data, weight=h.weight, ex_index=h.ex_index, **kw)

    @classmethod
    def _from_handle(cls, h):
        self = cls.__new__(cls)
        self.h = C.c_void_p(h)
        name = C.create_string_buffer(256)
        attrs, examples, width, nnz = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_uint64()
        _check(lib().dsb200_dataset_info(self.h, name, C.c_int(256), C.byref(attrs), C.byref(examples), C.byref(width), C.byref(nnz)))
        self.name, self.attributes, self.examples, self.width, self.nnz = name.value.decode(), attrs.value, examples.value, width.value, nnz.value
        return self

    def load_sparse(self, start, end, index, data=None):
        """NNDataSet::LoadSparseData: replace the contents (host arrays are copied and uploaded)."""
        _check(lib().dsb200_dataset_load_sparse(self.h, _p(start), _p(end), _p(index), _p(data)))

    def close(self):
        if self.h:
            lib().dsb200_dataset_destroy(self.h)
            self.h = C.c_void_p()


def load_netcdf(fname):
    """LoadNetCDF (E/NNTypes.cpp:2456): every dataset of a NetCDF classic file."""
    arr = (C.c_void_p * 16)()
    n = C.c_int()
    _check(lib().dsb200_datasets_load_netcdf(fname.encode(), arr, C.c_int(16), C.byref(n)))
    return [Dataset._from_handle(arr[i]) for i in range(n.value)]


def save_netcdf(fname, datasets):
    """SaveNetCDF: the datasets into one NetCDF (CDF-5) file."""
    _check(lib().dsb200_datasets_save_netcdf(fname.encode(), _handles(datasets), C.c_int(len(datasets))))


def _handles(datasets):
    arr = (C.c_void_p * max(len(datasets), 1))()
    for i, d in enumerate(datasets):
        arr[i] = d.h
    return arr


class Network:
    """NNNetwork: built from the reference's JSON layer-description language."""

    def __init__(self, json_text, batch, datasets):
        self.h = C.c_void_p()
        self.datasets = list(datasets)
        _check(lib().dsb200_network_load_json(C.byref(self.h), json_text.encode(), C.c_uint32(batch), _handles(datasets),
                                              C.c_int(len(datasets))))
        _check(lib().dsb200_network_load_datasets(self.h, _handles(datasets), C.c_int(len(datasets))))
        self.batch = batch

    @classmethod
    def from_netcdf(cls, fname, batch, datasets):
        """LoadNeuralNetworkNetCDF + LoadDataSets."""
        self = cls.__new__(cls)
        self.h = C.c_void_p()
        self.datasets = list(datasets)
        _check(lib().dsb200_network_load_netcdf(C.byref(self.h), fname.encode(), C.c_uint32(batch)))
        _check(lib().dsb200_network_load_datasets(self.h, _handles(datasets), C.c_int(len(datasets))))
        self.batch = batch
        return self

    def save_netcdf(self, fname):
        _check(lib().dsb200_network_save_netcdf(self.h, fname.encode()))

    def close(self):
        if self.h:
            lib().dsb200_network_destroy(self.h)
            self.h = C.c_void_p()

    def set_training_mode(self, mode):
        _check(lib().dsb200_network_set_training_mode(self.h, C.c_int(mode)))

    def set_fusion(self, flag):
        _check(lib().dsb200_network_set_fusion(self.h, C.c_int(int(flag))))

    def set_gemm_mode(self, mode):
        _check(lib().dsb200_network_set_gemm_mode(self.h, C.c_int(mode)))

    def set_shuffle_indices(self, flag):
        _check(lib().dsb200_network_set_shuffle_indices(self.h, C.c_int(int(flag))))

    def set_position(self, pos):
        _check(lib().dsb200_network_set_position(self.h, C.c_uint32(pos)))

    def get_shuffle_indices(self):
        n = C.c_uint32()
        _check(lib().dsb200_network_get_shuffle_indices(self.h, None, C.c_uint32(0), C.byref(n)))
        out = np.zeros(n.value, dtype=np.uint32)
        if n.value:
            _check(lib().dsb200_network_get_shuffle_indices(self.h, _p(out), C.c_uint32(out.size), C.byref(n)))
        return out

    def train(self, epochs, alpha, lam=0.0, lam1=0.0, mu=0.0, mu1=0.0):
        e = C.c_float()
        _check(lib().dsb200_network_train(self.h, C.c_uint32(epochs), C.c_float(alpha), C.c_float(lam), C.c_float(lam1),
                                          C.c_float(mu), C.c_float(mu1), C.byref(e)))
        return e.value

    def train_step(self, position, alpha, lam=0.0, lam1=0.0, mu=0.0, mu1=0.0):
        e = C.c_float()
        _check(lib().dsb200_network_train_step(self.h, C.c_uint32(position), C.c_float(alpha), C.c_float(lam), C.c_float(lam1),
                                               C.c_float(mu), C.c_float(mu1), C.byref(e)))
        return e.value

    def validate(self, samples=0):
        """NNNetwork::Validate: finite-difference gradient check through the training kernels; True when every probe passes"""
        ok = C.c_int(0)
        _check(lib().dsb200_network_validate(self.h, C.c_uint32(samples), C.byref(ok)))
        return bool(ok.value)

    def predict_batch(self):
        _check(lib().dsb200_network_predict_batch(self.h))

    def topk(self, layer, k, batch, filt=None):
        key = np.empty((batch, k), dtype=np.float32)
        val = np.empty((batch, k), dtype=np.uint32)
        _check(lib().dsb200_network_topk(self.h, layer.encode(), C.c_uint32(k), None if filt is None else filt.h, _p(key), _p(val)))
        return key, val

    def topk_global(self, layer, k, batch, filt=None):
        """Model parallel: top-K of the whole layer with global unit ids, the same on every rank."""
        key = np.empty((batch, k), dtype=np.float32)
        val = np.empty((batch, k), dtype=np.uint32)
        _check(lib().dsb200_network_topk_global(self.h, layer.encode(), C.c_uint32(k), None if filt is None else filt.h, _p(key), _p(val)))
        return key, val

    def set_weights(self, src, dst, W=None, b=None):
        W = None if W is None else np.ascontiguousarray(W, dtype=np.float32)
        b = None if b is None else np.ascontiguousarray(b, dtype=np.float32)
        _check(lib().dsb200_network_set_weights(self.h, src.encode(), dst.encode(), _p(W), C.c_uint64(0 if W is None else W.size),
                                                _p(b), C.c_uint64(0 if b is None else b.size)))

    def get_weights(self, src, dst):
        nW, nB = C.c_uint64(), C.c_uint64()
        _check(lib().dsb200_network_get_weights(self.h, src.encode(), dst.encode(), None, C.c_uint64(0), None, C.c_uint64(0),
                                                C.byref(nW), C.byref(nB)))
        W = np.empty(nW.value, dtype=np.float32)
        b = np.empty(nB.value, dtype=np.float32)
        _check(lib().dsb200_network_get_weights(self.h, src.encode(), dst.encode(), _p(W), nW, _p(b), nB, C.byref(nW), C.byref(nB)))
        return W, b

    def get_gradients(self, src, dst):
        n = C.c_uint64()
        _check(lib().dsb200_network_get_gradients(self.h, src.encode(), dst.encode(), None, C.c_uint64(0), C.byref(n)))
        g = np.empty(n.value, dtype=np.float32)
        _check(lib().dsb200_network_get_gradients(self.h, src.encode(), dst.encode(), _p(g), n, C.byref(n)))
        return g

    def _layer_buf(self, fn, layer):
        n = C.c_uint64()
        _check(fn(self.h, layer.encode(), None, C.c_uint64(0), C.byref(n)))
        out = np.empty(n.value, dtype=np.float32)
        _check(fn(self.h, layer.encode(), _p(out), n, C.byref(n)))
        return out

    def get_units(self, layer):
        return self._layer_buf(lib().dsb200_network_get_units, layer)

    def get_deltas(self, layer):
        return self._layer_buf(lib().dsb200_network_get_deltas, layer)

    def layer_info(self, layer):
        s, ls, a, b = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_uint32()
        _check(lib().dsb200_network_layer_info(self.h, layer.encode(), C.byref(s), C.byref(ls), C.byref(a), C.byref(b)))
        return s.value, ls.value, a.value, b.value


def autoencoder_json(hidden, error="ScaledMarginalCrossEntropy", smce=(1.0, 0.0, 1.0, 1.0), denoising_p=0.0,
                     sparseness=None, shuffle=False, in_name="gl_input", out_name="gl_output", activation="Sigmoid",
                     out_activation="Sigmoid", init=("Gaussian", 0.01, 0.0), p_dropout=0.0):
    """JSON in the reference's layer-description language for a sparse-in / sparse-out autoencoder
    (shape of samples/movielens/config.json and benchmarks/dsstne/config.json)."""
    import json
    wi = {"Scheme": init[0], "Scale": init[1], "Bias": init[2]}
    layers = [{"Name": "Input", "Kind": "Input", "N": "auto", "DataSet": in_name, "Sparse": True}]
    for i, n in enumerate(hidden):
        layers.append({"Name": f"Hidden{i + 1}", "Kind": "Hidden", "Type": "FullyConnected", "N": int(n), "Activation": activation,
                       "Sparse": bool(sparseness is not None), "WeightInit": wi})
        if p_dropout > 0:
            layers[-1]["pDropout"] = float(p_dropout)
    layers.append({"Name": "Output", "Kind": "Output", "Type": "FullyConnected", "DataSet": out_name, "N": "auto",
                   "Activation": out_activation, "Sparse": True, "WeightInit": wi})
    cfg = {"Version": 0.8, "Name": "AE", "Kind": "FeedForward", "ShuffleIndices": bool(shuffle),
           "ScaledMarginalCrossEntropy": {"oneTarget": smce[0], "zeroTarget": smce[1], "oneScale": smce[2], "zeroScale": smce[3]},
           "Layers": layers, "ErrorFunction": error}
    if denoising_p > 0:
        cfg["Denoising"] = {"p": denoising_p}
    if sparseness is not None:
        cfg["SparsenessPenalty"] = {"p": sparseness[0], "beta": sparseness[1]}
    return json.dumps(cfg)
