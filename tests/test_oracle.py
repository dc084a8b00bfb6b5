"""CPU: the oracle restatement against the golden vectors recorded from the reference's own code."""
import numpy as np
import torch

import golden_util as gu
from oracle import foldclass_oracle as orc


def _run_cases(db, lens, queries, cases):
    dbt, lt = torch.from_numpy(db), torch.from_numpy(lens.astype(np.float32))
    for c in cases:
        s, i, full = orc.search_torch_flavour(dbt, lt, torch.from_numpy(queries[c["qi"]]), c["qlen"], c["mincov"], c["k"])
        # same library, same formula: bit-identical scores; ids identical except inside exact ties
        np.testing.assert_array_equal(s.numpy(), c["scores"])
        orc.check_topk(s.numpy(), i.numpy(), c["scores"], c["ids"], full.numpy(), tol=0.0)


def test_torch_flavour_matches_reference_golden_n2048():
    _run_cases(*gu.torch_flavour_n2048())


def test_torch_flavour_matches_reference_golden_n300():
    _run_cases(*gu.torch_flavour_n300_full())


def test_config1_golden():
    db, lens, z = gu.config1()
    q = torch.from_numpy(z["query"][0])
    for mincov, tag in ((0.7, "mincov07"), (0.0, "mincov0")):
        s, i, full = orc.search_torch_flavour(torch.from_numpy(db), torch.from_numpy(lens.astype(np.float32)), q,
                                              int(z["qlen"]), mincov, int(z["k"]))
        np.testing.assert_array_equal(s.numpy(), z[f"scores_{tag}"])
        orc.check_topk(s.numpy(), i.numpy(), z[f"scores_{tag}"], z[f"ids_{tag}"], full.numpy(), tol=0.0)


def test_blockwise_ip_restatement_against_reference_on_unit_norm_db():
    """faiss flavour pinned indirectly: on a unit-norm DB, mask off, IP == the reference's cosine search."""
    db, z = gu.ip_flavour()
    xqn = orc.normalize_queries(torch.from_numpy(z["queries_raw"])).numpy()
    np.testing.assert_allclose(xqn, z["queries_normalised"], atol=1e-7)
    for bs in (262144, 8192, 1000):
        D, I = orc.knn_exact_blockwise(xqn, orc.db_iterator(db, bs), int(z["k"]))
        full = orc.all_scores_ip(xqn, db)
        for r in range(xqn.shape[0]):
            orc.check_topk(D[r], I[r], z["D"][r], z["I"][r], full[r], tol=1e-6)
    assert (I[:3, 0] == np.array([0, 66942, 31337])).all()  # planted neighbours


def test_blockwise_padding_and_threshold():
    db = np.eye(5, 128, dtype=np.float32)
    D, I = orc.knn_exact_blockwise(db[:2], orc.db_iterator(db, 2), 8)
    assert (I[:, 5:] == -1).all() and np.isneginf(D[:, 5:]).all()
    assert I[0, 0] == 0 and I[1, 0] == 1
    hi, hd, qi = orc.threshold_hits(D, I, 0.5)
    assert hi.tolist() == [0, 1] and qi.tolist() == [0, 1] and hd.tolist() == [1.0, 1.0]


def test_coverage_mask_is_fp32_product():
    lens = torch.arange(1, 2001, dtype=torch.float32)
    for mincov in (0.3, 0.6, 0.7, 1.0):
        for qlen in (15, 30, 150, 700):
            want = (qlen >= lens * mincov).float()
            got = orc.coverage_mask(qlen, lens, mincov)
            assert torch.equal(want, got)
            f32 = (np.float32(qlen) >= lens.numpy() * np.float32(mincov)).astype(np.float32)
            np.testing.assert_array_equal(got.numpy(), f32)
