"""CPU oracle for the Foldclass database-search hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, on the CPU, the arithmetic of the two search flavours of
psipred/merizo_search so that the CUDA path can be checked against it.  Only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl
reference`` legs of ``bench.py`` may import it.  The product package
(``merizo_search_b200``) never does: it fails loudly when the CUDA library is
missing.

Parity status:
  * torch flavour  -- PINNED: ``tests/golden/make_golden.py`` imports the
    reference's own ``search_query_against_db`` (dbsearch.py:75-81) in the build
    container and commits its outputs; ``tests/test_oracle.py`` checks this
    oracle against them.
  * faiss flavour  -- the arithmetic lives in the third-party ``faiss`` module
    (un-vendored, version unpinned: reference README.md:16, ansible
    roles/merizosearch/tasks/main.yml:29-35) which is not installable here.
    ``knn_exact_blockwise`` restates the published algorithm of
    ``IndexFlat(d, METRIC_INNER_PRODUCT)`` + ``ResultHeap`` (exact fp32 inner
    product, k largest per query, sorted descending) along the reference's own
    call sites (dbsearch.py:213-248).  It is pinned INDIRECTLY: on unit-norm
    databases with the coverage mask off, inner product == cosine, and the golden
    vectors of the reference's torch function cover that case.  Anything
    faiss-specific beyond that (tie order inside the heap) is "parity unpinned"
    and the comparison rule tolerates tie permutations.

All citations are relative to /root/reference/merizo_search/programs/Foldclass/.
"""
from __future__ import annotations

import json
from typing import Iterable, Iterator, Tuple

import numpy as np
import torch
import torch.nn.functional as F

DIM = 128


# --------------------------------------------------------------------------- #
# torch flavour  (in-memory ``.pt`` database)
# --------------------------------------------------------------------------- #
def coverage_mask(qlen: int, lengths: torch.Tensor, mincov: float) -> torch.Tensor:
    """dbsearch.py:76 -- ``(len(q.seq) >= lengths * mincov).float()``.

    ``lengths`` is fp32, so the product is an fp32 product (the python float is
    cast to fp32 by torch's scalar promotion) compared with the integer length.
    """
    return (qlen >= lengths * mincov).float()


def cosine_scores(db: torch.Tensor, q: torch.Tensor) -> torch.Tensor:
    """dbsearch.py:78 -- ``F.cosine_similarity(db[N,128], q[1,128], dim=-1)``.

    torch computes sum_d (x/max(|x|,1e-8)) * (y/max(|y|,1e-8)) in fp32.
    """
    return F.cosine_similarity(db, q.reshape(1, -1), dim=-1)


def search_torch_flavour(db: torch.Tensor, lengths: torch.Tensor, q_emb: torch.Tensor,
                         qlen: int, mincov: float, k: int) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """dbsearch.py:75-81.  Returns (top_scores[k], top_ids[k] int64, all_scores[N]).

    ``k > N`` raises (as ``torch.topk`` does).  Masked rows score exactly 0.
    """
    scores = cosine_scores(db, q_emb) * coverage_mask(qlen, lengths, mincov)
    top_scores, top_ids = torch.topk(scores, k, dim=0)
    return top_scores, top_ids, scores


# --------------------------------------------------------------------------- #
# faiss flavour  (memory-mapped ``.json`` database)
# --------------------------------------------------------------------------- #
def read_dbinfo(path: str) -> dict:
    """dbutil.py:24-25."""
    with open(path, "r") as fh:
        return json.load(fh)


def db_memmap(filename: str, shape: tuple) -> np.memmap:
    """dbutil.py:28-30 -- headerless fp32 C-order [DB_SIZE, DB_DIM]."""
    return np.memmap(filename, dtype="float32", mode="r", shape=shape)


def db_iterator(embeddings, batch_size: int) -> Iterator[np.ndarray]:
    """dbutil.py:33-35 -- consecutive row blocks."""
    for i0 in range(0, embeddings.shape[0], batch_size):
        yield embeddings[i0:i0 + batch_size]


def normalize_queries(xq: torch.Tensor) -> torch.Tensor:
    """dbsearch.py:303-304 -- ``F.normalize`` (p=2, dim=1, eps=1e-12)."""
    return F.normalize(xq)


def knn_exact_blockwise(xq, blocks: Iterable, k: int) -> Tuple[np.ndarray, np.ndarray]:
    """dbsearch.py:213-248 restated without faiss.

    Per block: exact fp32 inner products (IndexFlat.search), the block's k
    best, ``I += i0`` (dbsearch.py:238), merged into a running k-best per query
    (ResultHeap.add_result); ``finalize`` sorts each row by descending score.
    Rows never filled keep faiss's (-inf, -1) padding.
    """
    xq = torch.as_tensor(np.ascontiguousarray(xq), dtype=torch.float32)
    nq = xq.shape[0]
    best_d = torch.full((nq, k), float("-inf"), dtype=torch.float32)
    best_i = torch.full((nq, k), -1, dtype=torch.int64)
    i0 = 0
    for xb in blocks:
        xb = torch.as_tensor(np.ascontiguousarray(xb), dtype=torch.float32)
        ni = xb.shape[0]
        s = xq @ xb.T                                   # IndexFlat IP
        kk = min(k, ni)
        d, i = torch.topk(s, kk, dim=1)
        i = i + i0                                      # I += i0
        cat_d = torch.cat([best_d, d], dim=1)
        cat_i = torch.cat([best_i, i], dim=1)
        best_d, pos = torch.topk(cat_d, k, dim=1)       # heap merge + finalize order
        best_i = torch.gather(cat_i, 1, pos)
        i0 += ni
    return best_d.numpy(), best_i.numpy()


def threshold_hits(D: np.ndarray, I: np.ndarray, mincos: float):
    """dbsearch.py:318-326 -- query-major flattening of the hits >= mincos."""
    D_mask = np.where(D >= mincos)
    return I[D_mask], D[D_mask], D_mask[0]


def all_scores_ip(xq, xb) -> np.ndarray:
    """Full [nq, N] fp32 inner-product matrix (for the tie-tolerant comparison)."""
    xq = torch.as_tensor(np.ascontiguousarray(xq), dtype=torch.float32)
    xb = torch.as_tensor(np.ascontiguousarray(xb), dtype=torch.float32)
    return (xq @ xb.T).numpy()


# --------------------------------------------------------------------------- #
# comparison rule (north_star): same ids and order, |score diff| <= tol,
# ties within tolerance may permute.
# --------------------------------------------------------------------------- #
def check_topk(got_scores, got_ids, want_scores, want_ids, all_scores, tol: float = 1e-5,
               n_valid: int | None = None) -> None:
    """Raise AssertionError unless (got) is an acceptable answer given (want).

    ``all_scores`` is the oracle's full score vector for this query, used to
    decide whether an id mismatch is a tie within ``tol``.  ``n_valid`` = how
    many of the k slots are real (the rest must be (-inf, -1) padding).
    """
    got_scores = np.asarray(got_scores, dtype=np.float32).reshape(-1)
    got_ids = np.asarray(got_ids, dtype=np.int64).reshape(-1)
    want_scores = np.asarray(want_scores, dtype=np.float32).reshape(-1)
    want_ids = np.asarray(want_ids, dtype=np.int64).reshape(-1)
    all_scores = np.asarray(all_scores, dtype=np.float32).reshape(-1)
    k = want_ids.shape[0]
    assert got_ids.shape[0] == k and got_scores.shape[0] == k, "wrong k"
    nv = k if n_valid is None else n_valid
    for r in range(nv, k):
        assert got_ids[r] == -1 and np.isneginf(got_scores[r]), f"rank {r}: expected (-inf,-1) padding"
    g_ids, g_sc = got_ids[:nv], got_scores[:nv]
    assert len(set(g_ids.tolist())) == nv, "duplicate ids in result"
    assert (g_ids >= 0).all() and (g_ids < all_scores.shape[0]).all(), "id out of range"
    # reported score must be the true score of the reported id
    true_sc = all_scores[g_ids]
    err = np.abs(g_sc - true_sc)
    assert (err <= tol).all(), f"score/id mismatch: max err {err.max():.3e}"
    # rank-wise agreement with the oracle list, within tolerance
    derr = np.abs(g_sc - want_scores[:nv])
    assert (derr <= tol).all(), f"rank-wise score differs from oracle: max {derr.max():.3e}"
    # descending order (within tolerance)
    assert (np.diff(g_sc) <= tol).all(), "scores not sorted descending"
    # ids: equal, or a permutation among scores tied within tol
    mism = np.nonzero(g_ids != want_ids[:nv])[0]
    for r in mism:
        assert abs(float(all_scores[g_ids[r]]) - float(want_scores[r])) <= tol, (
            f"rank {r}: id {g_ids[r]} (score {all_scores[g_ids[r]]}) is not a tie of "
            f"oracle id {want_ids[r]} (score {want_scores[r]})")
