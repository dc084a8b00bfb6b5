"""GPU: the reference-facing Python layer (merizo_search_b200.dbsearch / faiss_driver) against the oracle."""
import pickle

import numpy as np
import pytest
import torch

from merizo_search_b200 import dbsearch as b200
from merizo_search_b200 import faiss_driver, synth
from oracle import foldclass_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture
def pt_db(tmp_path):
    """A `.pt` database as createdb writes it (makedb.py:85-91): raw [N,128] tensor + pickled index."""
    n = 3000
    rows = synth.host_db(n, base_seed=55, normalise=False) * 3.0
    lens = synth.host_lengths(n, seed=56)
    rng = np.random.default_rng(57)
    index = [(f"/db/dom{i:05d}.pdb", rng.standard_normal((int(l), 3)).astype(np.float32), "A" * int(l)) for i, l in enumerate(lens)]
    base = str(tmp_path / "toy")
    torch.save(torch.from_numpy(rows), base + ".pt")
    with open(base + ".index", "wb") as fh:
        pickle.dump(index, fh)
    yield base, rows, lens, index
    b200.release_all()


def test_read_database_and_search_pt_flavour(pt_db):
    base, rows, lens, index = pt_db
    target = b200.read_database(base, torch.device("cuda"))
    assert target["faiss"] is False and target["mdfn"] is None and target["mifn"] is None
    assert target["database"].size(0) == len(index) == len(target["index"])
    assert torch.equal(target["lengths"], torch.tensor([len(t[2]) for t in index], dtype=torch.float))
    again = b200.read_database(base, "cuda")  # multi_domain_search re-reads: same resident database
    assert again["database"] is target["database"]
    dbt, lt = torch.from_numpy(rows), torch.from_numpy(lens.astype(np.float32))
    for qi, (qlen, mincov, k) in enumerate([(150, 0.7, 10), (60, 0.7, 1), (400, 0.3, 50)]):
        emb = torch.from_numpy(synth.host_queries(1, 60 + qi))
        res = b200.search_query_against_db({"embedding": emb, "seq": "G" * qlen}, target, mincov, k)
        assert res["scores"].dtype == torch.float32 and res["indices"].dtype == torch.int64
        ws, wi, full = orc.search_torch_flavour(dbt, lt, emb[0], qlen, mincov, k)
        orc.check_topk(res["scores"].numpy(), res["indices"].numpy(), ws.numpy(), wi.numpy(), full.numpy(), tol=1e-5)
        name, coords, seq = target["index"][res["indices"][0]]  # consumers index python lists with the tensor element
        assert name.endswith(".pdb") and "{:.4f}".format(res["scores"][0])
    with pytest.raises(RuntimeError):
        b200.search_query_against_db({"embedding": emb, "seq": "G" * 100}, target, 0.7, len(index) + 1)


def test_knn_exact_block_iterator():
    db = synth.host_db(30000, base_seed=58)
    xq = orc.normalize_queries(torch.from_numpy(synth.host_queries(9, 59)))
    D, I = b200.knn_exact(xq, orc.db_iterator(db, 4096), 7)
    assert D.dtype == np.float32 and I.dtype == np.int64 and D.shape == (9, 7)
    wD, wI = orc.knn_exact_blockwise(xq.numpy(), orc.db_iterator(db, 4096), 7)
    full = orc.all_scores_ip(xq.numpy(), db)
    for r in range(9):
        orc.check_topk(D[r], I[r], wD[r], wI[r], full[r], tol=1e-5)
    b200.release_all()


def test_dbsearch_faiss_driver_skip_tmalign(tiny_faiss_db, tmp_path):
    d, emb, names, seqs, coords, metas = tiny_faiss_db
    proj = torch.from_numpy(np.random.default_rng(3).standard_normal((3, 128)).astype(np.float32))

    def network(x):  # stand-in embedder: any deterministic [1,L,3] -> [1,128] map
        return (x.mean(dim=1) @ proj.to(x.device)) + 0.1

    queries = []
    for j in (4, 17, 33):  # query = a database domain's own coordinates; its embedding is NOT the stored one
        queries.append({"coords": coords[j], "seq": seqs[j], "name": f"/q/query{j}.pdb", "dom_str": "1-10"})
    # plant the stand-in embeddings of the queries as rows of the database so that hits are certain
    dev = torch.device("cuda")
    q_emb = np.stack([network(torch.from_numpy(q["coords"]).unsqueeze(0)).numpy().reshape(-1) for q in queries])
    q_n = orc.normalize_queries(torch.from_numpy(q_emb)).numpy()
    emb2 = emb.copy()
    emb2[[4, 17, 33]] = q_n
    emb2.tofile(d / "t_raw_128d_norm.db")
    target = b200.read_database(str(d / "t"), dev)
    assert target == {"database": str(d / "t") + ".json", "faiss": True}
    results, all_results = faiss_driver.dbsearch_faiss(
        queries, target, str(tmp_path / "tmp"), network, topk=5, mincov=0.7, mincos=0.5, mintm=0.5, fastmode=True,
        device=dev, inputs_are_ca=True, skip_tmalign=True)
    D, I = orc.knn_exact_blockwise(q_n, orc.db_iterator(emb2, 16), 5)
    hi, hd, qi = orc.threshold_hits(D, I, 0.5)
    assert len(results) == int(qi.max()) + 1 == 3
    flat = [(q, r, h) for q, res in enumerate(results) for r, h in sorted(res.items())]
    assert len(flat) == len(hi)
    for (q, r, h), want_id, want_d, want_q in zip(flat, hi, hd, qi):
        assert q == want_q and int(h["dbindex"]) == int(want_id) and abs(float(h["score"]) - float(want_d)) <= 1e-5
        assert h["target"] == names[want_id] and h["t_len"] == len(seqs[want_id]) and h["metadata"] == metas[want_id]
        assert h["query"] == f"query{[4, 17, 33][q]}" and h["tmalign_output"] is None and h["dom_str"] == "1-10"
    assert results[0][0]["dbindex"] == 4 and results[1][0]["dbindex"] == 17 and results[2][0]["dbindex"] == 33
    b200.release_all()


def test_residency_server_keeps_the_pt_database_across_clients(pt_db, tmp_path, monkeypatch):
    """serve.py: a server process-alike holds the database in HBM; a client's read_database finds it
    through FCS_SERVER and its searches give the reference's answer without loading the matrix."""
    import threading

    from merizo_search_b200 import serve

    base, rows, lens, index = pt_db
    eng, info = serve._load(base, [0])
    sock = str(tmp_path / "fcs.sock")
    ready = threading.Event()
    th = threading.Thread(target=serve.serve, args=(eng, sock, info, ready), daemon=True)
    th.start()
    assert ready.wait(10)
    monkeypatch.setenv("FCS_SERVER", sock)
    b200.release_all()
    target = b200.read_database(base, torch.device("cuda"))
    assert isinstance(target["database"].engine, serve.RemoteEngine) and target["database"].size(0) == len(index)
    dbt, lt = torch.from_numpy(rows), torch.from_numpy(lens.astype(np.float32))
    emb = torch.from_numpy(synth.host_queries(1, 61))
    res = b200.search_query_against_db({"embedding": emb, "seq": "G" * 150}, target, 0.7, 10)
    ws, wi, full = orc.search_torch_flavour(dbt, lt, emb[0], 150, 0.7, 10)
    orc.check_topk(res["scores"].numpy(), res["indices"].numpy(), ws.numpy(), wi.numpy(), full.numpy(), tol=1e-5)
    target["database"].engine.shutdown_server()
    th.join(10)
    b200.release_all()
    eng.close()
