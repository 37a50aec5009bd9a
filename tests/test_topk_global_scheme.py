"""Model-parallel top-K (BASELINE config 5 at P = 8) -- the scheme of NNNetwork::CalculateTopKGlobal restated over
the CPU oracle: every rank takes the top-K of its column shard (exclusion filter rebased to the shard, as the engine shards
datasets, E/NNTypes.cpp:1864-1879), adds its first unit to the ids, the [batch][K] lists are laid side by side rank-major
(what dsb200_all_gather does with stride P * K) and one key/value top-K over the P * K candidates gives the result.  It must
equal the single-process top-K of the whole row BIT FOR BIT, ties included (descending score, ascending global id)."""
import numpy as np
import pytest

from helpers import ml20m


def shard_filter(h, lo, hi):
    """CSR restricted to columns [lo, hi), ids rebased to the shard."""
    start, end, index = [], [], []
    pos = 0
    for r in range(len(h.start)):
        row = h.index[int(h.start[r]):int(h.end[r])]
        row = row[(row >= lo) & (row < hi)] - lo
        start.append(pos)
        pos += len(row)
        end.append(pos)
        index.append(row)
    return (np.array(start, dtype=np.uint64), np.array(end, dtype=np.uint64),
            np.concatenate(index).astype(np.uint32) if index else np.zeros(0, np.uint32))


@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("ties", [False, True], ids=["distinct", "heavy-ties"])
@pytest.mark.parametrize("filtered", [False, True], ids=["plain", "filtered"])
def test_sharded_topk_then_merge_equals_global_topk(orc, dsb, world, ties, filtered):
    rng = np.random.default_rng(100 * world + ties)
    B, N, K = 24, 5003, 40                                            # N not divisible by any world size: uneven shards
    scores = (rng.integers(0, 30, size=(B, N)) if ties else rng.permutation(B * N).reshape(B, N)).astype(np.float32)
    h = ml20m(examples=B, width=N) if filtered else None
    want_k, want_v = orc.topk(scores, K, filt=None if h is None else (h.start, h.end, h.index))
    keys, vals = [], []
    for r in range(world):
        lo, hi = orc.shard_range(N, r, world)
        filt = None if h is None else shard_filter(h, lo, hi)
        k_loc, v_loc = orc.topk(np.ascontiguousarray(scores[:, lo:hi]), K, filt=filt)
        keys.append(k_loc)
        vals.append((v_loc + np.uint32(lo)).astype(np.uint32))        # dsb200_topk_offset
    full_k = np.ascontiguousarray(np.concatenate(keys, axis=1))       # [B][world * K], rank-major: dsb200_all_gather
    full_v = np.ascontiguousarray(np.concatenate(vals, axis=1))
    got_k, got_v = orc.topk(full_k, K, value=full_v)                  # dsb200_topk_kv
    np.testing.assert_array_equal(got_k, want_k)
    np.testing.assert_array_equal(got_v, want_v)


def test_shards_smaller_than_k_leave_sentinels_that_never_win(orc, dsb):
    rng = np.random.default_rng(5)
    B, N, K, world = 4, 50, 20, 8                                     # 6-7 columns per rank < K: local lists end in (-MAX_VALUE, 0)
    scores = rng.standard_normal((B, N)).astype(np.float32)
    want_k, want_v = orc.topk(scores, K)
    keys, vals = [], []
    for r in range(world):
        lo, hi = orc.shard_range(N, r, world)
        k_loc, v_loc = orc.topk(np.ascontiguousarray(scores[:, lo:hi]), K)
        keys.append(k_loc)
        vals.append((v_loc + np.uint32(lo)).astype(np.uint32))
    got_k, got_v = orc.topk(np.ascontiguousarray(np.concatenate(keys, axis=1)), K, value=np.ascontiguousarray(np.concatenate(vals, axis=1)))
    np.testing.assert_array_equal(got_k, want_k)
    np.testing.assert_array_equal(got_v, want_v)
