"""CPU: host logic of the reference-facing layer that needs no GPU -- module splicing, error behaviour, the
faiss-flavour driver (hit flattening, record retrieval, result dicts) with the oracle standing in for the engine."""
import types

import numpy as np
import pytest
import torch

from merizo_search_b200 import dbsearch as b200
from merizo_search_b200 import faiss_driver, native
from oracle import foldclass_oracle as orc


def test_install_patches_reference_module():
    ref = types.SimpleNamespace(read_database=None, search_query_against_db=None, dbsearch_faiss=None)
    out = b200.install(ref)
    assert out.read_database is b200.read_database and out.search_query_against_db is b200.search_query_against_db
    assert out.dbsearch_faiss is faiss_driver.dbsearch_faiss


def test_read_database_missing_logs_and_exits():
    with pytest.raises(SystemExit) as e:  # reference: logger.error + sys.exit(1) (dbsearch.py:70-72)
        b200.read_database("/nonexistent/db", "cuda")
    assert e.value.code == 1


def test_read_database_json_flavour_returns_path_only(tiny_faiss_db):
    d = tiny_faiss_db[0]
    assert b200.read_database(str(d / "t"), "cuda") == {"database": str(d / "t") + ".json", "faiss": True}


def test_device_list_from_env(monkeypatch):
    monkeypatch.setenv("FCS_DEVICES", "2, 5,7")
    assert b200._devices() == [2, 5, 7]
    monkeypatch.delenv("FCS_DEVICES")
    assert b200._devices() is None


def test_search_query_rejects_foreign_database():
    with pytest.raises(TypeError):
        b200.search_query_against_db({"embedding": torch.zeros(1, 128), "seq": "AAA"}, {"database": torch.zeros(3, 128)}, 0.7, 1)


class _OracleEngine:
    """Stands in for engine.LocalEngine: same search() contract, arithmetic by the CPU oracle."""

    def __init__(self, rows):
        self.rows, self.n_rows = rows, rows.shape[0]

    def search(self, q, k, qlen=None, mincov=0.0, qnorm=native.QNORM_NONE, mode=native.MODE_AUTO, kprime=0):
        q = torch.from_numpy(np.ascontiguousarray(q, dtype=np.float32))
        if qnorm == native.QNORM_L2:
            q = orc.normalize_queries(q)
        return orc.knn_exact_blockwise(q.numpy(), orc.db_iterator(self.rows, 16), k)


def test_dbsearch_faiss_driver_logic_with_oracle_engine(tiny_faiss_db, tmp_path, monkeypatch):
    d, emb, names, seqs, coords, metas = tiny_faiss_db
    import os

    key = (os.path.abspath(str(d / "t_raw_128d_norm.db")), "faiss")
    monkeypatch.setitem(b200._RESIDENT, key, b200.ResidentDatabase(_OracleEngine(emb), "faiss"))

    def network(x):  # stand-in embedder: query j embeds exactly onto database row j (x 3: un-normalised)
        j = int(round(float(x[0, 0, 0])))
        return torch.from_numpy(emb[j:j + 1] * 3.0)

    queries = []
    for j in (4, 17):
        c = coords[j].copy()
        c[0, 0] = j
        queries.append({"coords": c, "seq": seqs[j], "name": f"/q/query{j}.pdb"})
    results, all_results = faiss_driver.dbsearch_faiss(
        queries, {"database": str(d / "t") + ".json", "faiss": True}, str(tmp_path / "tmp"), network, topk=3, mincov=0.7,
        mincos=0.2, mintm=0.5, fastmode=True, device=torch.device("cpu"), inputs_are_ca=True, skip_tmalign=True)
    xq = orc.normalize_queries(torch.from_numpy(np.stack([emb[4] * 3, emb[17] * 3]))).numpy()
    D, I = orc.knn_exact_blockwise(xq, orc.db_iterator(emb, 16), 3)
    hi, hd, qi = orc.threshold_hits(D, I, 0.2)
    flat = [(q, h) for q, res in enumerate(results) for _, h in sorted(res.items())]
    assert len(flat) == len(hi) >= 2
    for (q, h), want_id, want_d, want_q in zip(flat, hi, hd, qi):
        assert q == want_q and int(h["dbindex"]) == int(want_id) and abs(float(h["score"]) - float(want_d)) < 1e-6
        assert h["target"] == names[want_id] and h["t_len"] == len(seqs[want_id]) and h["metadata"] == metas[want_id]
        assert h["q_len"] == len(queries[q]["seq"]) and h["tmalign_output"] is None and h["dom_str"] is None
    assert results[0][0]["dbindex"] == 4 and results[1][0]["dbindex"] == 17
    assert "{:.4f}".format(results[0][0]["score"]) == "1.0000"
    # no hit above the threshold: empty result instead of the reference's np.max([]) crash (documented divergence)
    r2, a2 = faiss_driver.dbsearch_faiss(queries, {"database": str(d / "t") + ".json", "faiss": True}, str(tmp_path / "tmp"),
                                         network, topk=3, mincov=0.7, mincos=1.5, mintm=0.5, fastmode=True,
                                         device=torch.device("cpu"), inputs_are_ca=True, skip_tmalign=True)
    assert r2 == [] and a2 == []


def test_faiss_flavour_refuses_a_non_cuda_device(tiny_faiss_db):
    """-d cpu must not silently put the database on the GPUs (nor fall back to a CPU path that does not exist)."""
    d, emb, *_ = tiny_faiss_db
    b200._RESIDENT.clear()
    with pytest.raises(native.FcsError):
        b200.load_resident_file(str(d / "t_raw_128d_norm.db"), emb.shape[0], torch.device("cpu"))
    with pytest.raises(native.FcsError):  # a file too short for the advertised row count
        b200.load_resident_file(str(d / "t_raw_128d_norm.db"), emb.shape[0] + 1, "cuda")
