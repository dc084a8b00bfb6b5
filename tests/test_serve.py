"""CPU: the residency server's protocol and the client hook (serve.py) -- the engine behind the server is the CPU oracle
here (the protocol does not care); on a GPU box the same server wraps engine.LocalEngine."""
import os
import threading

import numpy as np
import pytest
import torch

from merizo_search_b200 import dbsearch as b200
from merizo_search_b200 import native, serve, synth
from oracle import foldclass_oracle as orc


class _OracleEngine:
    def __init__(self, rows):
        self.rows, self.n_rows, self.n_shards = rows, rows.shape[0], 1

    def search(self, q, k, qlen=None, mincov=0.0, qnorm=native.QNORM_NONE, mode=native.MODE_AUTO, kprime=0):
        q = torch.from_numpy(np.ascontiguousarray(q, dtype=np.float32))
        if qnorm == native.QNORM_L2:
            q = orc.normalize_queries(q)
        if k > self.n_rows + 5:
            raise ValueError("k far too large")
        return orc.knn_exact_blockwise(q.numpy(), orc.db_iterator(self.rows, 64), k)

    def close(self):
        pass


@pytest.fixture
def server(tmp_path):
    db = synth.host_db(500, base_seed=5)
    path = tmp_path / "x_raw_128d_norm.db"
    db.tofile(path)
    sock = str(tmp_path / "fcs.sock")
    info = {"path": os.path.abspath(str(path)), "flavour": "faiss", "stamp": serve.file_stamp(str(path)), "n_rows": 500, "shards": 1}
    ready = threading.Event()
    th = threading.Thread(target=serve.serve, args=(_OracleEngine(db), sock, info, ready), daemon=True)
    th.start()
    assert ready.wait(10)
    yield db, str(path), sock, th
    try:
        c = serve.RemoteEngine(sock)
        c.shutdown_server()
        c.close()
    except Exception:
        pass
    th.join(5)


def test_remote_engine_answers_like_the_engine_behind_it(server):
    db, path, sock, _ = server
    eng = serve.RemoteEngine(sock)
    assert eng.n_rows == 500 and eng.holds(path, "faiss") and not eng.holds(path, "pt")
    q = synth.host_queries(7, 3)
    s, i = eng.search(q, 10, qnorm=native.QNORM_L2)
    D, I = orc.knn_exact_blockwise(orc.normalize_queries(torch.from_numpy(q)).numpy(), orc.db_iterator(db, 64), 10)
    np.testing.assert_array_equal(i, I)
    np.testing.assert_allclose(s, D, atol=0)
    with pytest.raises(native.FcsError):  # a server-side failure comes back as an error, and the server keeps serving
        eng.search(q, 10_000)
    s2, _ = eng.search(q[:1], 3)
    assert s2.shape == (1, 3)
    eng.close()


def test_drop_in_loader_uses_the_server_only_for_the_same_unchanged_file(server, tmp_path, monkeypatch):
    db, path, sock, _ = server
    monkeypatch.setenv("FCS_SERVER", sock)
    b200._RESIDENT.clear()
    resident = b200.load_resident_file(path, 500)
    assert isinstance(resident.engine, serve.RemoteEngine) and resident.size(0) == 500
    D, I = resident.engine.search(db[:3], 1)
    assert I[:, 0].tolist() == [0, 1, 2]
    b200._RESIDENT.clear()
    # another file (or the same file rewritten) is NOT served from the stale resident copy
    other = tmp_path / "y_raw_128d_norm.db"
    db[::-1].copy().tofile(other)
    assert serve.connect(str(other), "faiss") is None
    os.utime(path, ns=(1, 1))
    assert serve.connect(path, "faiss") is None
    monkeypatch.delenv("FCS_SERVER")
    assert serve.connect(path, "faiss") is None
