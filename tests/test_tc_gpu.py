"""GPU: the tensor-core batched path (K3 tcgen05 GEMM + filter, select rounds, K4 fp32 rescore)."""
import numpy as np
import pytest
import torch

from merizo_search_b200 import native, synth
from oracle import foldclass_oracle as orc

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _db(rows, **kw):
    h = native.Database(rows.shape[0], keep_bf16=True, **kw)
    h.upload(0, rows)
    h.finalize()
    return h


@pytest.mark.parametrize("n,nq", [(64, 128), (100, 7), (4096, 600), (1000, 513)])
def test_tcgen05_scores_match_bf16_matmul(n, nq):
    """Raw accumulators of the UMMA pipeline against a bf16-rounded fp32 matmul (descriptor/layout check)."""
    db = synth.host_db(n, base_seed=21)
    q = synth.host_queries(nq, batch_id=21, normalise=True)
    h = _db(db)
    got = h.debug_tc_approx(q)
    want = (torch.from_numpy(q).bfloat16().float() @ torch.from_numpy(db).bfloat16().float().T).numpy()
    assert not np.isnan(got).any(), f"{np.isnan(got).sum()} (query,row) pairs were never produced"
    np.testing.assert_allclose(got, want, atol=2e-5, rtol=0)
    h.close()


@pytest.mark.parametrize("n,nq,k", [(70001, 700, 10), (70001, 33, 100), (300000, 1500, 50), (5000, 40, 128)])
def test_tc_search_is_exact(n, nq, k):
    db = synth.host_db(n, base_seed=31)
    q = synth.host_queries(nq, batch_id=31, planted_from=db, planted_ids=np.arange(0, n, max(1, n // 16))[:16])
    h = _db(db, id_offset=5)
    s, i = h.search(q, k, qnorm=native.QNORM_L2, mode=native.MODE_TC)
    t = h.timing()
    assert t.last_mode == native.MODE_TC
    xqn = orc.normalize_queries(torch.from_numpy(q)).numpy()
    D, I = orc.knn_exact_blockwise(xqn, orc.db_iterator(db, 262144), k)
    full = orc.all_scores_ip(xqn, db)
    for r in range(nq):
        ids = i[r].copy()
        ids[ids >= 0] -= 5
        orc.check_topk(s[r], ids, D[r], I[r], full[r], tol=TOL, n_valid=min(n, k))
    assert t.last_tc_fallbacks <= max(2, nq // 50), f"{t.last_tc_fallbacks} certificate fallbacks on iid data"
    h.close()


@pytest.mark.parametrize("n,nq,k", [(200000, 96, 129), (200000, 70, 500), (400000, 40, 2048), (60000, 33, 1000)])
def test_tc_search_large_k_is_exact(n, nq, k):
    """k above the register-resident list size (128) and up to FCS_MAX_K: larger candidate buffers, the block-sort rescore,
    multi-pass exact scan for the fallback queue.  `-k` is unbounded in the reference (merizo.py:134)."""
    db = synth.host_db(n, base_seed=33)
    q = synth.host_queries(nq, batch_id=33, planted_from=db, planted_ids=np.arange(0, n, max(1, n // 8))[:8])
    h = _db(db)
    s, i = h.search(q, k, qnorm=native.QNORM_L2, mode=native.MODE_TC)
    t = h.timing()
    assert t.last_mode == native.MODE_TC
    xqn = orc.normalize_queries(torch.from_numpy(q)).numpy()
    D, I = orc.knn_exact_blockwise(xqn, orc.db_iterator(db, 262144), k)
    full = orc.all_scores_ip(xqn, db)
    for r in range(nq):
        orc.check_topk(s[r], i[r], D[r], I[r], full[r], tol=TOL, n_valid=min(n, k))
    assert t.last_tc_fallbacks <= max(2, nq // 10), f"{t.last_tc_fallbacks} certificate fallbacks on iid data"
    # AUTO must keep a large batch with a large k on the tensor-core path
    h.search(synth.host_queries(512, 1, normalise=True), k)
    assert h.timing().last_mode == native.MODE_TC
    h.close()


def test_tc_matches_gemv_path_bitwise_ids():
    n, nq, k = 50000, 64, 10
    db = synth.host_db(n, base_seed=41)
    q = synth.host_queries(nq, batch_id=41, normalise=True)
    h = _db(db)
    s1, i1 = h.search(q, k, mode=native.MODE_TC)
    s2, i2 = h.search(q, k, mode=native.MODE_GEMV)
    np.testing.assert_array_equal(i1, i2)
    np.testing.assert_allclose(s1, s2, atol=2e-6)
    h.close()


def test_certificate_failure_falls_back_to_exact_scan():
    """Thousands of near-duplicates of the query: bf16 cannot separate them, the certificate must fail
    and the exact fp32 scan must take over."""
    n, k = 20000, 10
    rng = np.random.Generator(np.random.PCG64(3))
    db = synth.host_db(n, base_seed=51)
    base = db[123].copy()
    dup = base[None, :] + 2e-4 * rng.standard_normal((6000, 128)).astype(np.float32)
    db[5000:11000] = dup / np.linalg.norm(dup, axis=1, keepdims=True)
    # 20 queries inside the duplicate cluster (more than one 8-query fallback pass), 20 ordinary ones
    q = np.stack([db[5000 + 37 * j] for j in range(20)] + [synth.host_queries(1, 52 + j, normalise=True)[0] for j in range(20)])
    q = np.ascontiguousarray(q, dtype=np.float32)
    h = _db(db)
    s, i = h.search(q, k, mode=native.MODE_TC)
    assert h.timing().last_tc_fallbacks >= 10
    D, I = orc.knn_exact_blockwise(q, orc.db_iterator(db, 262144), k)
    full = orc.all_scores_ip(q, db)
    for r in range(q.shape[0]):
        orc.check_topk(s[r], i[r], D[r], I[r], full[r], tol=TOL)
    h.close()


def test_candidate_overflow_falls_back_to_exact_scan():
    """Rows ordered by increasing similarity: every round floods the candidate buffer."""
    n, k = 40000, 10
    db = synth.host_db(n, base_seed=61)
    q = synth.host_queries(40, batch_id=61, normalise=True)
    order = np.argsort(db @ q[0])  # ascending similarity to query 0
    db = np.ascontiguousarray(db[order])
    h = _db(db)
    s, i = h.search(q, k, mode=native.MODE_TC)
    D, I = orc.knn_exact_blockwise(q, orc.db_iterator(db, 262144), k)
    full = orc.all_scores_ip(q, db)
    for r in range(q.shape[0]):
        orc.check_topk(s[r], i[r], D[r], I[r], full[r], tol=TOL)
    h.close()


def test_auto_mode_picks_paths():
    """AUTO compares a cost estimate of the two paths: small batches and small databases scan, big ones use tcgen05."""
    db = synth.host_db(300000, base_seed=71)
    h = _db(db)
    h.search(synth.host_queries(2, 1, normalise=True), 5)
    assert h.timing().last_mode == native.MODE_GEMV
    h.search(synth.host_queries(512, 1, normalise=True), 5)
    assert h.timing().last_mode == native.MODE_TC
    h.close()
    h = _db(db)
    h.search(synth.host_queries(8, 1, normalise=True), 5)  # one scan pass always beats the TC path's fixed cost
    assert h.timing().last_mode == native.MODE_GEMV
    h.search(synth.host_queries(16, 1, normalise=True), 5)  # two passes over 300 k rows vs 125 us + one sweep: scan still wins
    assert h.timing().last_mode == native.MODE_GEMV
    h.search(synth.host_queries(64, 1, normalise=True), 5)
    assert h.timing().last_mode == native.MODE_TC
    h.search(synth.host_queries(64, 1, normalise=True), 5, qlen=None)
    h.close()
