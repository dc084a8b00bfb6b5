"""CPU: the converted layout of a .pt database's .index file (merizo_search_b200/dbindex.py) answers exactly what the
reference asks of the unpickled list (dbsearch.py:53-58, 124, 157; dbsearch_fulllength.py:381-462)."""
import os
import pickle

import numpy as np
import pytest
import torch

from merizo_search_b200 import dbindex


def _make_index(path, n=37, seed=4):
    """The reference's layout: list[(pdb path, ca float32 [L,3], seq str)] pickled to <db>.index (makedb.py:68-91)."""
    rng = np.random.default_rng(seed)
    entries = []
    for i in range(n):
        L = int(rng.integers(1, 60))
        seq = "".join(rng.choice(list("ACDEFGHIKLMNPQRSTVWYX"), size=L))
        entries.append((f"/data/cath/dompdb/{i:07d}é.pdb" if i == 5 else f"/data/cath/dompdb/{i:07d}.pdb",
                        rng.standard_normal((L, 3)).astype(np.float32), seq))
    with open(path + ".index", "wb") as fh:
        pickle.dump(entries, fh)
    return entries


def _same(a, b):
    return a[0] == b[0] and a[2] == b[2] and a[1].dtype == np.float32 and np.array_equal(a[1], b[1])


def test_roundtrip_and_list_protocol(tmp_path):
    base = str(tmp_path / "db")
    entries = _make_index(base)
    out = dbindex.convert_index(base)
    assert out == base + ".index.fcs" and sorted(os.listdir(out)) == sorted([a + ".npy" for a in dbindex._ARRAYS] + ["meta.json"])
    lazy = dbindex.open_if_fresh(base)
    assert lazy is not None and len(lazy) == len(entries)
    for i, e in enumerate(entries):
        assert _same(lazy[i], e)
    assert all(_same(a, b) for a, b in zip(lazy, entries))          # iteration
    assert _same(lazy[-1], entries[-1]) and _same(lazy[np.int64(3)], entries[3])
    assert _same(lazy[torch.tensor([7, 2])[0]], entries[7])          # dbsearch.py:124 indexes with a tensor element
    assert [x[0] for x in lazy[2:5]] == [e[0] for e in entries[2:5]]
    with pytest.raises(IndexError):
        lazy[len(entries)]
    name, coords, seq = lazy[0]
    coords[:] = 0  # a private copy: the mapped file is not writable through the record
    assert np.array_equal(lazy[0][1], entries[0][1])
    assert np.array_equal(lazy.lengths(), np.asarray([len(e[2]) for e in entries], dtype=np.int32))
    assert lazy.names([4, 5]) == [entries[4][0], entries[5][0]]


def test_stale_or_missing_sidecar_falls_back_to_the_pickle(tmp_path, monkeypatch):
    base = str(tmp_path / "db")
    entries = _make_index(base)
    monkeypatch.delenv("FCS_INDEX_CACHE", raising=False)
    idx, lengths = dbindex.load_index(base)                          # no sidecar: the reference's own path
    assert isinstance(idx, list) and len(idx) == len(entries) and not os.path.exists(base + ".index.fcs")
    assert np.array_equal(lengths, [len(e[2]) for e in entries])
    monkeypatch.setenv("FCS_INDEX_CACHE", "1")
    dbindex.load_index(base)                                         # slow load once, sidecar written
    idx2, lengths2 = dbindex.load_index(base)
    assert isinstance(idx2, dbindex.LazyIndex) and np.array_equal(lengths2, lengths)
    _make_index(base, n=12, seed=9)                                  # the pickle is rebuilt: the sidecar is stale
    assert dbindex.open_if_fresh(base) is None
    idx3, lengths3 = dbindex.load_index(base)                        # ... and is replaced by a fresh one
    assert len(idx3) == 12 and len(lengths3) == 12
    assert isinstance(dbindex.load_index(base)[0], dbindex.LazyIndex)


def test_corrupt_sidecar_is_ignored(tmp_path):
    base = str(tmp_path / "db")
    _make_index(base)
    d = dbindex.convert_index(base)
    with open(os.path.join(d, "meta.json"), "w") as fh:
        fh.write("{not json")
    assert dbindex.open_if_fresh(base) is None


def test_empty_index(tmp_path):
    base = str(tmp_path / "db")
    with open(base + ".index", "wb") as fh:
        pickle.dump([], fh)
    lazy = dbindex.LazyIndex(dbindex.convert_index(base))
    assert len(lazy) == 0 and list(lazy) == [] and lazy.lengths().shape == (0,)
