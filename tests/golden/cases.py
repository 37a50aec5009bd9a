"""Seeded input builders shared by the golden-vector generator (run once on the GPU box against the REFERENCE's own
CUDA kernels, oracle/_ref/libdsstne_refkernels.so) and by the CPU-side test that checks the oracle against the
stored reference outputs.  Small shapes so the fixtures stay a few hundred KB."""
import numpy as np

BATCH, WIDTH, STRIDE = 64, 512, 128


def make_csr(seed, examples=BATCH, width=WIDTH, mean=12, weighted=False, analog=False):
    rng = np.random.Generator(np.random.PCG64(seed))
    lens = np.clip(rng.binomial(width, mean / width, size=examples), 0, width)
    lens[0] = 0                                            # one empty row
    rows = [np.sort(rng.choice(width, size=int(n), replace=False)).astype(np.uint32) for n in lens]
    end = np.cumsum([len(r) for r in rows]).astype(np.uint64)
    start = np.concatenate([[0], end[:-1]]).astype(np.uint64)
    index = np.concatenate(rows).astype(np.uint32)
    data = rng.uniform(0.5, 5.0, size=index.size).astype(np.float32) if analog else None
    weight = rng.uniform(0.5, 1.5, size=examples).astype(np.float32) if weighted else None
    return dict(start=start, end=end, index=index, data=data, weight=weight, width=width)


def dense_inputs(seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    return dict(W=(rng.standard_normal((WIDTH, STRIDE)) * 0.05).astype(np.float32),
                Z0=rng.standard_normal((BATCH, STRIDE)).astype(np.float32),
                z_out=(rng.standard_normal((BATCH, WIDTH)) * 2.0 - 1.0).astype(np.float32),
                delta=(rng.standard_normal((BATCH, STRIDE)) * 0.1).astype(np.float32),
                rnd=rng.random(4096).astype(np.float32),
                g=(rng.standard_normal(WIDTH * 8) * 0.01).astype(np.float32),
                w=(rng.standard_normal(WIDTH * 8) * 0.05).astype(np.float32),
                v=(rng.random(WIDTH * 8) * 0.01).astype(np.float32),
                gv=(rng.random(WIDTH * 8) * 0.01).astype(np.float32),
                scores=rng.permutation(BATCH * 4096).astype(np.float32).reshape(BATCH, 4096))


Z_CASES = [("boolean", {}), ("weighted", dict(weighted=True)), ("analog", dict(weighted=True, analog=True)),
           ("denoised", dict(denoised=True)), ("analog_denoised", dict(weighted=True, analog=True, denoised=True))]
LOSS_CASES = [(3, 0, False), (3, 0, True), (2, 0, False), (1, 0, False), (1, 3, False), (2, 7, False)]   # (error, activation, ignoreZero)
OPT_HP = dict(alpha=0.025, lam=1e-4, lam1=1e-5, mu=0.9, mu1=0.999, t=3.0)
SMCE = (0.8, 0.05, 1.5, 0.75)
BOOST = (2.0, 0.5)
DENOISE_P = 0.2
