"""Synthetic sparse datasets in DSSTNE's CSR-with-start/end layout (E/NNTypes.h:213-225).

Shapes follow SURVEY.md section 8d: C1 rows ~ Binomial(N, 1%) nnz; C2..C5 rows ~ log-normal
fitted to MovieLens-20M (mean 144.4, min 20, max 9,254) with columns drawn from a
Zipf-Mandelbrot popularity law p_i ~ 1/(i+q) over a random permutation of the items
(q=30 puts the most popular item in ~49% of rows at ML-20M density, as in ML-20M).
Column ids are unique and ascending inside a row (what generateNetCDF emits for
one sample line, U/NetCDFhelper.cpp:281-330).  Seeds: data 12134 (the reference's
FIXED_SEED, U/Utils.h:27), weights 12345.
"""
import numpy as np

DATA_SEED = 12134
WEIGHT_SEED = 12345


class HostCsr:
    """Plain host arrays; `examples` rows over `width` columns."""

    def __init__(self, start, end, index, width, data=None, weight=None, ex_index=None):
        self.start = np.ascontiguousarray(start, dtype=np.uint64)
        self.end = np.ascontiguousarray(end, dtype=np.uint64)
        self.index = np.ascontiguousarray(index, dtype=np.uint32)
        self.width = int(width)
        self.data = data
        self.weight = weight
        self.ex_index = ex_index

    @property
    def examples(self):
        return len(self.ex_index) if self.ex_index is not None else len(self.start)

    @property
    def nnz(self):
        return int(self.index.shape[0])

    def row_lengths(self):
        return (self.end - self.start).astype(np.int64)


def _row_lengths(rng, examples, width, dist, mean):
    if dist == "binomial":
        n = rng.binomial(width, mean / width, size=examples)
        return np.clip(n, 1, width)
    if dist == "lognormal":
        sigma = 1.0
        mu = np.log(mean) - 0.5 * sigma * sigma
        n = np.rint(rng.lognormal(mu, sigma, size=examples)).astype(np.int64)
        return np.clip(n, 20, min(9254, width))
    if dist == "fixed":
        return np.full(examples, int(mean), dtype=np.int64)
    raise ValueError(dist)


def make_csr(examples, width, mean_nnz, dist="lognormal", col="zipf", seed=DATA_SEED, zipf_q=30.0,
             analog=False, weighted=False, empty_rows=0):
    """Returns a HostCsr.  `empty_rows` forces that many rows to have no non-zeros."""
    rng = np.random.Generator(np.random.PCG64(seed))
    lens = _row_lengths(rng, examples, width, dist, mean_nnz)
    if empty_rows:
        lens[rng.choice(examples, size=empty_rows, replace=False)] = 0
    perm = rng.permutation(width).astype(np.uint32)
    if col == "zipf":
        p = 1.0 / (np.arange(width, dtype=np.float64) + 1.0 + zipf_q)
        cdf = np.cumsum(p)
        cdf /= cdf[-1]
    rows = []
    for n in lens:
        n = int(n)
        if n == 0:
            rows.append(np.empty(0, dtype=np.uint32))
            continue
        if col == "uniform" or n > width // 2:
            ids = rng.choice(width, size=n, replace=False).astype(np.uint32)
        else:
            got = np.empty(0, dtype=np.int64)
            while got.size < n:
                draw = np.searchsorted(cdf, rng.random(int((n - got.size) * 1.5) + 16))
                cand = np.concatenate([got, draw])
                _, first = np.unique(cand, return_index=True)
                got = cand[np.sort(first)]          # keep arrival order, drop repeats
            ids = perm[got[:n]]
        rows.append(np.sort(ids).astype(np.uint32))
    start = np.zeros(examples, dtype=np.uint64)
    end = np.zeros(examples, dtype=np.uint64)
    pos = 0
    for i, r in enumerate(rows):
        start[i] = pos
        pos += r.size
        end[i] = pos
    index = np.concatenate(rows) if rows else np.empty(0, dtype=np.uint32)
    data = rng.uniform(0.5, 5.0, size=index.size).astype(np.float32) if analog else None
    weight = rng.uniform(0.5, 1.5, size=examples).astype(np.float32) if weighted else None
    return HostCsr(start, end, index, width, data=data, weight=weight)


def make_weights(sizes, seed=WEIGHT_SEED, scale=0.01, out_bias=0.0):
    """Gaussian N(0, scale) weights W[l]: [sizes[l]][sizes[l+1]] and zero biases
    (benchmarks/dsstne/config.json:17 uses Gaussian init); injected, never cuRAND."""
    rng = np.random.Generator(np.random.PCG64(seed))
    Ws, bs = [], []
    for i in range(len(sizes) - 1):
        Ws.append((rng.standard_normal((sizes[i], sizes[i + 1])) * scale).astype(np.float32))
        b = np.zeros(sizes[i + 1], dtype=np.float32)
        if i == len(sizes) - 2 and out_bias != 0.0:
            b[:] = out_bias
        bs.append(b)
    return Ws, bs
