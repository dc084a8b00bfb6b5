#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the REFERENCE's own code.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

It imports ``search_query_against_db`` (dbsearch.py:75-81), ``read_pdb``
(Foldclass/utils.py:42) and ``FoldClassNet`` (nndef_fold_egnn_embed.py:34) from
/root/reference and records their outputs on seeded inputs.  The fixtures are
committed; nothing in tests/ reads /root/reference at run time.

Databases larger than a few hundred rows are not stored: they are regenerated
from ``merizo_search_b200.synth`` seeds, and a sha256 of the bytes is stored so a
drifting RNG is detected instead of silently changing the inputs.
"""
import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/merizo_search")

from programs.Foldclass.dbsearch import search_query_against_db  # noqa: E402
from programs.Foldclass.nndef_fold_egnn_embed import FoldClassNet  # noqa: E402
from programs.Foldclass.utils import read_pdb  # noqa: E402

from merizo_search_b200 import synth  # noqa: E402

REF_DB = "/root/reference/examples/database"


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def ref_search(db, lengths, q, qlen, mincov, k):
    r = search_query_against_db(
        {"embedding": torch.from_numpy(q).reshape(1, -1), "seq": "A" * int(qlen)},
        {"database": torch.from_numpy(db), "lengths": torch.from_numpy(lengths.astype(np.float32))},
        mincov, k)
    return r["scores"].numpy().astype(np.float32), r["indices"].numpy().astype(np.int64)


def ted_lengths(n):
    idx = np.fromfile(os.path.join(REF_DB, "ted100_9606_small/ted100_9606_small_seq.index"), dtype=np.int64)
    idx = idx.reshape(-1, 2)
    return (idx[:n, 1] - idx[:n, 0]).astype(np.int32)


def main():
    torch.manual_seed(0)
    torch.set_num_threads(1)  # deterministic reduction order in the reference run

    # ---- A. torch flavour, raw (un-normalised) rows, real TED lengths -------------------
    N = 2048
    db = synth.host_db(N, base_seed=101, normalise=False)
    # vary the row norms over a few orders of magnitude (the .pt DB is raw network output)
    scale = np.exp(np.random.Generator(np.random.PCG64(5)).normal(0, 1.5, size=(N, 1))).astype(np.float32)
    db = (db * scale).astype(np.float32)
    db[17] = 0.0  # a zero row: cosine_similarity clamps the norm at 1e-8 -> score 0
    lengths = ted_lengths(N)
    queries = synth.host_queries(6, batch_id=1, planted_from=db, planted_ids=np.array([3, 500, 2047]))
    cases = []
    for qi, qlen, mincov, k in [
        (0, 150, 0.7, 1), (0, 150, 0.7, 10), (0, 150, 0.7, 100),
        (1, 150, 0.0, 10), (2, 30, 0.7, 10), (2, 30, 0.7, 100), (3, 15, 0.3, 10), (3, 15, 0.6, 50),
        (4, 683, 0.7, 10), (5, 90, 0.7, 128), (4, 2000, 1.0, 2048),
    ]:
        s, i = ref_search(db, lengths, queries[qi], qlen, mincov, k)
        cases.append(dict(qi=qi, qlen=qlen, mincov=mincov, k=k, scores=s, ids=i))
    np.savez_compressed(
        os.path.join(HERE, "torch_flavour_n2048.npz"),
        db_seed=101, db_sha=sha(db), scale_seed=5, lengths=lengths, queries=queries,
        n_cases=len(cases),
        **{f"c{j}_{key}": np.asarray(val) for j, c in enumerate(cases) for key, val in c.items()})
    print("torch_flavour_n2048:", len(cases), "cases; db sha", sha(db)[:12])

    # ---- A2. a tiny database stored in full (no RNG dependence at all) ------------------
    Ns = 300
    dbs = synth.host_db(Ns, base_seed=202, normalise=False)
    lens_s = ted_lengths(Ns)
    qs = synth.host_queries(3, batch_id=2)
    small = []
    for qi, qlen, mincov, k in [(0, 120, 0.7, 10), (1, 60, 0.7, 300), (2, 400, 0.5, 1)]:
        s, i = ref_search(dbs, lens_s, qs[qi], qlen, mincov, k)
        small.append(dict(qi=qi, qlen=qlen, mincov=mincov, k=k, scores=s, ids=i))
    np.savez_compressed(
        os.path.join(HERE, "torch_flavour_n300_full.npz"), db=dbs, lengths=lens_s, queries=qs,
        n_cases=len(small),
        **{f"c{j}_{key}": np.asarray(val) for j, c in enumerate(small) for key, val in c.items()})
    print("torch_flavour_n300_full:", len(small), "cases")

    # ---- B. BASELINE config 1: M0.pdb query vs a DB of the bundled CATH size ------------
    # Weights/embeddings are missing from the reference checkout (.MISSING_LARGE_BLOBS), so
    # the embedder runs with seeded random weights and the DB is synthetic.
    torch.manual_seed(1234)
    net = FoldClassNet(128).eval()
    qd = read_pdb(pdbfile="/root/reference/examples/M0.pdb", pdb_chain="A")
    with torch.no_grad():
        emb = net(torch.from_numpy(qd["coords"]).unsqueeze(0)).numpy().astype(np.float32)
    Nc = 14942
    dbc = synth.host_db(Nc, base_seed=303, normalise=False)
    lens_c = synth.host_lengths(Nc, seed=9)
    qlen = len(qd["seq"])
    s, i = ref_search(dbc, lens_c, emb[0], qlen, 0.7, 10)
    s0, i0 = ref_search(dbc, lens_c, emb[0], qlen, 0.0, 10)
    np.savez_compressed(
        os.path.join(HERE, "config1_m0_vs_cath_size.npz"), db_seed=303, db_sha=sha(dbc), len_seed=9,
        lengths_sha=sha(lens_c), query=emb, qlen=qlen, k=10,
        scores_mincov07=s, ids_mincov07=i, scores_mincov0=s0, ids_mincov0=i0)
    print("config1: qlen", qlen, "top ids", i[:5], "scores", s[:3])

    # ---- C. unit-norm DB of the bundled TED-slice size, mask off: cosine == inner product
    Nt = 66943
    dbt = synth.host_db(Nt, base_seed=404, normalise=True)
    ones = np.ones(Nt, dtype=np.int32)
    xq = synth.host_queries(8, batch_id=3, planted_from=dbt, planted_ids=np.array([0, 66942, 31337]))
    xqn = torch.nn.functional.normalize(torch.from_numpy(xq)).numpy()  # dbsearch.py:303-304
    D = np.zeros((8, 10), np.float32)
    I = np.zeros((8, 10), np.int64)
    for r in range(8):
        D[r], I[r] = ref_search(dbt, ones, xqn[r], 1, 0.0, 10)
    np.savez_compressed(
        os.path.join(HERE, "ip_flavour_n66943.npz"), db_seed=404, db_sha=sha(dbt), queries_raw=xq,
        queries_normalised=xqn, k=10, D=D, I=I)
    print("ip_flavour_n66943: top-1 ids", I[:, 0])


if __name__ == "__main__":
    main()
