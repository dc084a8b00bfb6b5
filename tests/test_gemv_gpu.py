"""GPU: the fp32 streaming kernel (K2) through the C ABI against the oracle and the reference goldens."""
import numpy as np
import pytest
import torch

import golden_util as gu
from merizo_search_b200 import native, synth
from oracle import foldclass_oracle as orc

pytestmark = pytest.mark.gpu
TOL = 1e-5  # north_star: cosine scores within 1e-5 absolute


def _torch_flavour_db(db, lens):
    h = native.Database(db.shape[0], normalise_rows=True, has_lengths=True)
    h.upload(0, db, lens)
    h.finalize()
    return h


def _check_torch_cases(db, lens, queries, cases):
    h = _torch_flavour_db(db, lens)
    dbt, lt = torch.from_numpy(db), torch.from_numpy(lens.astype(np.float32))
    for c in cases:
        s, i = h.search(queries[c["qi"]], c["k"], qlen=np.array([c["qlen"]]), mincov=c["mincov"],
                        qnorm=native.QNORM_COSINE, mode=native.MODE_GEMV)
        _, _, full = orc.search_torch_flavour(dbt, lt, torch.from_numpy(queries[c["qi"]]), c["qlen"], c["mincov"], c["k"])
        orc.check_topk(s[0], i[0], c["scores"], c["ids"], full.numpy(), tol=TOL)
    h.close()


def test_golden_torch_flavour_n2048():
    _check_torch_cases(*gu.torch_flavour_n2048())


def test_golden_torch_flavour_n300():
    _check_torch_cases(*gu.torch_flavour_n300_full())


def test_golden_config1():
    db, lens, z = gu.config1()
    h = _torch_flavour_db(db, lens)
    dbt, lt = torch.from_numpy(db), torch.from_numpy(lens.astype(np.float32))
    for mincov, tag in ((0.7, "mincov07"), (0.0, "mincov0")):
        s, i = h.search(z["query"], int(z["k"]), qlen=np.array([int(z["qlen"])]), mincov=mincov,
                        qnorm=native.QNORM_COSINE, mode=native.MODE_GEMV)
        _, _, full = orc.search_torch_flavour(dbt, lt, torch.from_numpy(z["query"][0]), int(z["qlen"]), mincov, 10)
        orc.check_topk(s[0], i[0], z[f"scores_{tag}"], z[f"ids_{tag}"], full.numpy(), tol=TOL)
    h.close()


def test_golden_ip_flavour_batches():
    db, z = gu.ip_flavour()
    h = native.Database(db.shape[0])
    for r0 in range(0, db.shape[0], 20000):  # block-wise feed, like db_iterator
        h.upload(r0, db[r0:r0 + 20000])
    h.finalize()
    full = orc.all_scores_ip(z["queries_normalised"], db)
    for nq in (1, 2, 3, 4, 5, 8):  # exercises the 1/2/4-query instantiations and the group loop
        s, i = h.search(z["queries_raw"][:nq], 10, qnorm=native.QNORM_L2, mode=native.MODE_GEMV)
        for r in range(nq):
            orc.check_topk(s[r], i[r], z["D"][r], z["I"][r], full[r], tol=TOL)
    h.close()


@pytest.mark.parametrize("n,k", [(1, 1), (5, 10), (31, 31), (33, 7), (1000, 64), (4097, 100), (4097, 128),
                                 (5000, 300), (3000, 2048), (70001, 10)])
def test_ragged_sizes_and_k(n, k):
    db = synth.host_db(n, base_seed=11 + n)
    q = synth.host_queries(2, batch_id=n, normalise=True)
    h = native.Database(n, id_offset=1000)
    h.upload(0, db)
    h.finalize()
    s, i = h.search(q, k, mode=native.MODE_GEMV)
    D, I = orc.knn_exact_blockwise(q, orc.db_iterator(db, 262144), k)
    full = orc.all_scores_ip(q, db)
    for r in range(2):
        ids = i[r].copy()
        ids[ids >= 0] -= 1000  # global ids carry the shard offset (I += i0, dbsearch.py:238)
        orc.check_topk(s[r], ids, D[r], I[r], full[r], tol=TOL, n_valid=min(n, k))
    h.close()


def test_many_query_groups_back_to_back():
    """nq > 8 = several launches in a row on one stream (they overlap through programmatic dependent launch and
    share the handle's scratch): every group must still be exact."""
    n, nq, k = 60000, 130, 10
    db = synth.host_db(n, base_seed=15)
    q = synth.host_queries(nq, 15, normalise=True)
    h = native.Database(n)
    h.upload(0, db)
    h.finalize()
    D, I = orc.knn_exact_blockwise(q, orc.db_iterator(db, 262144), k)
    full = orc.all_scores_ip(q, db)
    for rep in range(3):
        s, i = h.search(q, k, mode=native.MODE_GEMV)
        for r in range(nq):
            orc.check_topk(s[r], i[r], D[r], I[r], full[r], tol=TOL)
    h.close()


def test_exact_ties_are_ordered_by_ascending_id():
    """Duplicate rows tie exactly; torch/faiss leave the order unspecified, this library returns ascending ids."""
    n, k = 5000, 12
    db = synth.host_db(n, base_seed=19)
    dup = [4000, 17, 2500, 999, 3]
    for r in dup[1:]:
        db[r] = db[dup[0]]
    q = db[dup[0]].copy()
    for keep_bf16, mode in ((False, native.MODE_GEMV), (True, native.MODE_TC)):
        h = native.Database(n, keep_bf16=keep_bf16)
        h.upload(0, db)
        h.finalize()
        qs = np.stack([q] + [synth.host_queries(1, 90 + j, normalise=True)[0] for j in range(39)]).astype(np.float32)
        s, i = h.search(qs, k, mode=mode)
        assert i[0, :5].tolist() == sorted(dup) and np.allclose(s[0, :5], 1.0, atol=1e-6)
        D, I = orc.knn_exact_blockwise(qs, orc.db_iterator(db, 262144), k)
        full = orc.all_scores_ip(qs, db)
        for r in range(qs.shape[0]):
            orc.check_topk(s[r], i[r], D[r], I[r], full[r], tol=TOL)
        h.close()


def test_all_rows_masked_gives_zero_scores():
    n = 2000
    db = synth.host_db(n, base_seed=5, normalise=False)
    lens = np.full(n, 500, np.int32)
    h = _torch_flavour_db(db, lens)
    s, i = h.search(synth.host_queries(1, 9), 10, qlen=np.array([10]), mincov=0.7, qnorm=native.QNORM_COSINE,
                    mode=native.MODE_GEMV)
    assert (s == 0).all() and len(set(i[0].tolist())) == 10
    h.close()


def test_errors_are_codes_not_crashes():
    with pytest.raises(native.FcsError):
        native.Database(0)
    h = native.Database(10)
    with pytest.raises(native.FcsError):  # search before finalize
        h.search(np.zeros((1, 128), np.float32), 1)
    with pytest.raises(native.FcsError):  # finalize before all rows are there
        h.finalize()
    h.upload(0, np.zeros((10, 128), np.float32))
    h.finalize()
    with pytest.raises(native.FcsError):
        h.search(np.zeros((1, 128), np.float32), 0)
    with pytest.raises(native.FcsError):
        h.search(np.zeros((1, 128), np.float32), 4096)
    h.close()
