"""CPU: host-side logic -- shard partition, packed keys, host merge, faiss-flavour file readers."""
import json
import os

import numpy as np
import pytest
import torch

from merizo_search_b200 import engine, faiss_driver, synth
from host_merge import merge_keys_host
from oracle import foldclass_oracle as orc


def test_shard_ranges_cover_and_match_reference_offsets():
    for n, g in [(10, 3), (365_000_000, 8), (7, 8), (0, 2), (500_000, 1), (4097, 4)]:
        r = engine.shard_ranges(n, g)
        assert len(r) == g and r[0][0] == 0 and r[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
        per = -(-n // g) if n else 0
        assert all(hi - lo <= per for lo, hi in r)
    assert engine.shard_ranges(365_000_000, 8)[7] == (319_375_000, 365_000_000)


def test_key_roundtrip_and_order():
    rng = np.random.default_rng(0)
    s = rng.standard_normal(1000).astype(np.float32)
    s[:6] = [0.0, -0.0, np.inf, -np.inf, 1e-38, -1e-38]
    ids = rng.permutation(1000).astype(np.int64)
    k = engine.encode_keys(s, ids)
    s2, i2 = engine.decode_keys(k)
    np.testing.assert_array_equal(s.view(np.uint32), s2.view(np.uint32))
    np.testing.assert_array_equal(ids, i2)
    order = np.argsort(-k.astype(np.float64), kind="stable")  # monotone enough for a sanity check on distinct scores
    assert (np.diff(s[np.argsort(k)[::-1]]) <= 0).all()       # bigger key <=> bigger-or-equal score
    # ties: lower id wins
    kk = engine.encode_keys(np.array([0.5, 0.5], np.float32), np.array([7, 3]))
    assert kk[1] > kk[0]
    # padding
    e = engine.encode_keys(np.array([-np.inf], np.float32), np.array([-1]))
    assert e[0] == 0 and engine.decode_keys(e)[1][0] == -1 and np.isneginf(engine.decode_keys(e)[0][0])


def test_host_merge_equals_oracle_topk():
    db = synth.host_db(9000, base_seed=3)
    q = synth.host_queries(5, 3, normalise=True)
    k = 12
    parts = []
    for lo, hi in engine.shard_ranges(9000, 4):
        D, I = orc.knn_exact_blockwise(q, orc.db_iterator(db[lo:hi], 1000), k)
        parts.append(engine.encode_keys(D, I + lo))
    s, i = merge_keys_host(np.stack(parts), k)
    D, I = orc.knn_exact_blockwise(q, orc.db_iterator(db, 262144), k)
    full = orc.all_scores_ip(q, db)
    for r in range(5):
        orc.check_topk(s[r], i[r], D[r], I[r], full[r], tol=1e-6)


def test_threshold_hits_matches_reference_flattening():
    rng = np.random.default_rng(1)
    D = np.sort(rng.random((6, 5)).astype(np.float32), axis=1)[:, ::-1]
    I = rng.integers(0, 100, (6, 5))
    a = faiss_driver.threshold_hits(D, I, 0.5)
    b = orc.threshold_hits(D, I, 0.5)
    for x, y in zip(a, b):
        np.testing.assert_array_equal(x, y)


def test_record_files_vectorised_readers(tiny_faiss_db):
    d, emb, names, seqs, coords, metas = tiny_faiss_db
    info = faiss_driver.read_dbinfo(str(d / "t.json"))
    assert info == orc.read_dbinfo(str(d / "t.json"))
    mm = faiss_driver.embedding_memmap(str(d / info["dbfname_IP"]), info["DB_SIZE"], info["DB_DIM"])
    np.testing.assert_array_equal(np.asarray(mm), emb)
    blocks = list(faiss_driver.row_blocks(mm, 16))
    ref_blocks = list(orc.db_iterator(orc.db_memmap(str(d / info["dbfname_IP"]), (50, 128)), 16))
    assert [b.shape for b in blocks] == [b.shape for b in ref_blocks] == [(16, 128)] * 3 + [(2, 128)]
    rec = faiss_driver.RecordFiles(str(d), info)
    ids = np.array([49, 0, 7, 7, 23])
    assert rec.names(ids) == [names[i] for i in ids]
    assert rec.sequences(ids) == [seqs[i] for i in ids]
    assert rec.lengths(ids).tolist() == [len(seqs[i]) for i in ids]  # from the offset table alone
    assert rec.metadata(ids) == [metas[i] for i in ids]
    for got, i in zip(rec.coords(ids), ids):
        np.testing.assert_array_equal(got, coords[i])
    assert rec.has_metadata()


def test_plan_shards_small_databases_stay_on_one_gpu():
    """CATH scale (500 k rows) must not be spread over 8 GPUs for a 44 us scan; TED scale needs all of them."""
    assert engine.plan_shards(14_942, 8) == 1
    assert engine.plan_shards(500_000, 8) == 1
    assert engine.plan_shards(10_000_000, 8) == 2
    assert engine.plan_shards(365_000_000, 8) == 8
    assert engine.plan_shards(365_000_000, 8, bytes_per_row=768, free_bytes=170e9) == 8
    # memory forces more shards than the size rule asks for
    assert engine.plan_shards(6_000_000, 8, bytes_per_row=768, free_bytes=2e9) == 3
    with pytest.raises(Exception):
        engine.plan_shards(365_000_000, 1, bytes_per_row=768, free_bytes=170e9)


def test_tc_round_plan_visits_every_tile_exactly_once():
    """The tensor-core path's sampled rounds + complement sweep partition the shard's tiles (host mirror of tile_of)."""
    from merizo_search_b200 import native

    for n_rows in (1, 100, 4096, 4097, 5000, 8191, 8192, 70001, 300000, 1_250_000, 3_333_333):
        for kp, nq in ((64, 1024), (160, 4096), (512, 65536)):
            plan = native.debug_tc_plan(n_rows, kp, nq)
            nt = (n_rows + 127) // 128
            seen = []
            for r in plan:
                seen += [native.debug_tc_tile_of(r["j0"], r["stride"], r["comp_t"], i) for i in range(r["tiles"])]
            assert sorted(seen) == list(range(nt)), (n_rows, kp, plan)
            assert plan[0]["first"] == 1 and plan[0]["tiles"] * 128 <= 4096 and plan[-1]["partition"] == 1 and plan[-1]["rank"] == kp
            assert all(1 <= r["rank"] <= 1024 for r in plan[:-1])
