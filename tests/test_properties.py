"""CPU property tests (hypothesis) of the host-side invariants the multi-GPU plumbing rests on."""
import numpy as np
from hypothesis import given, settings
from hypothesis import strategies as st

from host_merge import merge_keys_host
from merizo_search_b200 import engine

finite_f32 = st.floats(width=32, allow_nan=False, allow_infinity=False)


@settings(max_examples=200, deadline=None)
@given(st.lists(st.tuples(finite_f32, st.integers(0, 2**32 - 2)), min_size=2, max_size=40))
def test_packed_keys_order_like_score_desc_then_id_asc(pairs):
    scores = np.asarray([p[0] for p in pairs], dtype=np.float32)
    ids = np.asarray([p[1] for p in pairs], dtype=np.int64)
    keys = engine.encode_keys(scores, ids)
    s2, i2 = engine.decode_keys(keys)
    assert np.array_equal(s2.view(np.uint32), scores.view(np.uint32)) and np.array_equal(i2, ids)  # bit-exact round trip
    order = np.argsort(keys)[::-1]
    # -0.0 and +0.0 are distinct keys (-0.0 below +0.0); compare with that total order
    ref = sorted(range(len(pairs)), key=lambda j: (-float(scores[j]), np.signbit(scores[j]), int(ids[j])))
    assert [(float(scores[j]), int(ids[j])) for j in order] == [(float(scores[j]), int(ids[j])) for j in ref]


@settings(max_examples=200, deadline=None)
@given(st.integers(0, 10**7), st.integers(1, 16))
def test_shard_ranges_partition_the_rows(n, g):
    r = engine.shard_ranges(n, g)
    assert len(r) == g and r[0][0] == 0 and r[-1][1] == n
    assert all(a[1] == b[0] for a, b in zip(r, r[1:])) and all(lo <= hi for lo, hi in r)
    sizes = [hi - lo for lo, hi in r]
    assert max(sizes) == -(-n // g) if n else max(sizes) == 0          # ceil(N/G) rows per shard, like I += i0 blocks


@settings(max_examples=200, deadline=None)
@given(st.lists(st.integers(1, 3000), min_size=0, max_size=200), st.integers(1, 9))
def test_embed_partition_is_a_contiguous_cover(lens, parts):
    p = engine.embed_partition(lens, parts)
    assert len(p) == parts and p[0][0] == 0 and p[-1][1] == len(lens)
    assert all(a[1] == b[0] for a, b in zip(p, p[1:])) and all(lo <= hi for lo, hi in p)
    if lens:
        cost = [sum(l * l for l in lens[lo:hi]) for lo, hi in p]
        # no slice exceeds its fair share by more than one structure's cost
        assert max(cost) <= sum(cost) / parts + max(l * l for l in lens) + 1e-9


@settings(max_examples=100, deadline=None)
@given(st.integers(1, 4), st.integers(1, 5), st.integers(1, 12), st.integers(0, 2**31 - 1))
def test_merge_of_sorted_lists_is_topk_of_the_union(n_lists, nq, k, seed):
    rng = np.random.default_rng(seed)
    scores = rng.standard_normal((n_lists, nq, k)).astype(np.float32)
    ids = rng.permutation(n_lists * nq * k).reshape(n_lists, nq, k).astype(np.int64)  # unique ids
    keys = np.sort(engine.encode_keys(scores, ids), axis=2)[:, :, ::-1]  # each list sorted, best first
    s, i = merge_keys_host(np.ascontiguousarray(keys), k)
    allk = np.sort(keys.transpose(1, 0, 2).reshape(nq, -1), axis=1)[:, ::-1][:, :k]
    ws, wi = engine.decode_keys(allk)
    assert np.array_equal(np.asarray(i), wi) and np.array_equal(np.asarray(s), ws)
