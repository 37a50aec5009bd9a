"""Generates tests/golden/ref_topksort.npz with the REFERENCE's own CPU top-K comparator, topKsort<float,uint32>
(U/Utils.cpp:213-243, compiled unmodified into oracle/_ref/libdsstne_refutils.so), on the (batch, K, N) shapes of the
reference's tst/gputests/TestSort.cpp:209-211.  Runs on the CPU (this container)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
SHAPES = [(8, 128, 1024), (4, 128, 100000), (8, 64, 1024), (8, 32, 64), (8, 1, 64)]     # batch reduced 128 -> 8/4 to keep the fixture small


def keys_for(i, B, N):
    rng = np.random.Generator(np.random.PCG64(12345 + i))
    return rng.permutation(B * N).astype(np.float32).reshape(B, N)          # tie-free, like the reference test's rand() data


if __name__ == "__main__":
    from oracle import oracle as orc
    out = {}
    for i, (B, K, N) in enumerate(SHAPES):
        k, v = orc.ref_topksort(keys_for(i, B, N), K)
        out[f"key_{i}"], out[f"val_{i}"] = k, v
    np.savez_compressed(os.path.join(HERE, "ref_topksort.npz"), **out)
    print("ok")
