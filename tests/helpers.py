"""Shared helpers for the parity tests: one synthetic dataset -> oracle view + device view."""
import numpy as np

from dsstne_b200 import datagen


def to_oracle(orc, h, random=None, ex_index=None):
    return orc.Csr(h.start, h.end, h.index, data=h.data, weight=h.weight,
                   ex_index=ex_index if ex_index is not None else h.ex_index, random=random)


def to_device(dsb, h, random=None, ex_index=None):
    return dsb.DeviceCsr(h.start, h.end, h.index, data=h.data, weight=h.weight,
                         ex_index=ex_index if ex_index is not None else h.ex_index, random=random)


def tiny(examples=256, width=2048, mean=20.5, **kw):
    """BASELINE.json config 1: 2,048 items, ~1% density, batch 256."""
    return datagen.make_csr(examples, width, mean, dist="binomial", col="uniform", **kw)


def ml20m(examples=1024, width=27278, mean=144.4, **kw):
    """BASELINE.json config 2 shape: MovieLens-20M-like rows / column popularity."""
    return datagen.make_csr(examples, width, mean, dist="lognormal", col="zipf", **kw)


def with_long_rows(h, lens, seed=7):
    """Replace the first len(lens) rows by rows of the given lengths (exercises the >4,608-nnz
    chunk loop of the reference kernels, E/kernels.cu:675-679, and our split-row path)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    rows = [h.index[int(s):int(e)] for s, e in zip(h.start, h.end)]
    for i, n in enumerate(lens):
        rows[i] = np.sort(rng.choice(h.width, size=n, replace=False)).astype(np.uint32)
    start = np.zeros(len(rows), dtype=np.uint64)
    end = np.zeros(len(rows), dtype=np.uint64)
    pos = 0
    for i, r in enumerate(rows):
        start[i] = pos
        pos += len(r)
        end[i] = pos
    out = datagen.HostCsr(start, end, np.concatenate(rows), h.width)
    if h.data is not None:
        out.data = rng.uniform(0.5, 5.0, size=out.nnz).astype(np.float32)
    out.weight = h.weight
    return out


def rel_err(a, b):
    """Relative error used by every fp32 parity assertion:  max |a-b| / (|b| + rms(b)).

    The tensor's own RMS is the floor for entries that are (near) zero through cancellation --
    fp32 rounding error scales with the magnitude of the terms summed, not with the magnitude of
    a result that happens to cancel; a plain |a-b|/|b| is unbounded there for ANY two correct
    fp32 implementations (the CPU oracle and the reference's own CUDA kernels differ by more
    than 1e-5 under it)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape
    rms = float(np.sqrt(np.mean(b * b))) if b.size else 0.0
    return float((np.abs(a - b) / (np.abs(b) + max(rms, 1e-30))).max()) if b.size else 0.0


def canon_columns(tstart, tend, tindex, tdata=None):
    """Per-column sorted (row, data) lists -- the canonical form of a transposed matrix."""
    out = []
    for s, e in zip(tstart, tend):
        rows = tindex[s:e]
        if tdata is None:
            out.append(np.sort(rows))
        else:
            o = np.lexsort((tdata[s:e], rows))
            out.append((rows[o], tdata[s:e][o]))
    return out
