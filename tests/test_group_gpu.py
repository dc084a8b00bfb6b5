"""GPU: the shard group (csrc/fcs_group.cu) behind LocalEngine -- several row shards driven by one host thread, key
lists exchanged device-to-device and merged on the first GPU.  Shards may share a device, so the whole exchange + merge
path runs on a single-GPU box too; with >= 2 GPUs the same tests also run with one shard per device."""
import numpy as np
import pytest
import torch

from merizo_search_b200 import engine, native, synth
from oracle import foldclass_oracle as orc

pytestmark = pytest.mark.gpu


def _device_sets():
    sets = [[0, 0, 0]]
    if torch.cuda.is_available() and torch.cuda.device_count() >= 2:
        sets.append(list(range(min(4, torch.cuda.device_count()))))
    return sets


def _check(s, i, q, db, k):
    D, I = orc.knn_exact_blockwise(q, orc.db_iterator(db, 262144), k)
    full = orc.all_scores_ip(q, db)
    for r in range(q.shape[0]):
        orc.check_topk(s[r], i[r], D[r], I[r], full[r], tol=1e-5, n_valid=min(db.shape[0], k))


@pytest.mark.parametrize("devices", _device_sets())
@pytest.mark.parametrize("n,nq,k,mode", [(50001, 3, 10, native.MODE_GEMV), (120000, 200, 20, native.MODE_TC),
                                          (30000, 64, 100, native.MODE_AUTO), (1000, 5, 10, native.MODE_GEMV)])
def test_group_equals_oracle(devices, n, nq, k, mode):
    db = synth.host_db(n, base_seed=81)
    q = synth.host_queries(nq, 81, normalise=True)
    eng = engine.LocalEngine(n, devices=devices, keep_bf16=True)
    assert eng.n_shards == len(devices) and eng.ranges == engine.shard_ranges(n, len(devices))
    eng.upload_blocks(orc.db_iterator(db, 7777))  # blocks straddle the shard boundaries
    eng.finalize()
    s, i = eng.search(q, k, mode=mode)
    _check(s, i, q, db, k)
    eng.close()


def test_group_torch_flavour_mask_and_normalisation():
    """.pt flavour through the group: raw rows normalised on the device, coverage mask, cosine -- vs the reference arithmetic."""
    n, k = 30000, 10
    db = synth.host_db(n, base_seed=83, normalise=False)
    lens = synth.host_lengths(n)
    q = synth.host_queries(1, 83)
    eng = engine.LocalEngine(n, devices=[0, 0], normalise_rows=True, has_lengths=True)
    eng.upload(0, db, lens)
    eng.finalize()
    s, i = eng.search(q, k, qlen=np.array([150]), mincov=0.7, qnorm=native.QNORM_COSINE, mode=native.MODE_GEMV)
    ws, wi, full = orc.search_torch_flavour(torch.from_numpy(db), torch.from_numpy(lens.astype(np.float32)), torch.from_numpy(q[0]), 150, 0.7, k)
    orc.check_topk(s[0], i[0], ws.numpy(), wi.numpy(), full.numpy(), tol=1e-5)
    eng.close()


def test_group_upload_file_reads_each_shards_byte_range(tmp_path):
    n, k = 70000, 10
    db = synth.host_db(n, base_seed=85)
    path = tmp_path / "x_raw_128d_norm.db"
    with open(path, "wb") as fh:
        fh.write(b"\0" * 4096)  # a header the loader must skip (file_offset)
        db.tofile(fh)
    q = synth.host_queries(40, 85, normalise=True)
    eng = engine.LocalEngine(n, devices=[0, 0, 0], keep_bf16=True)
    eng.upload_file(str(path), file_offset=4096)
    eng.finalize()
    s, i = eng.search(q, k, mode=native.MODE_TC)
    _check(s, i, q, db, k)
    eng.close()
    # a file that is too short is an error, not garbage
    eng = engine.LocalEngine(n, devices=[0, 0])
    with pytest.raises(native.FcsError):
        eng.upload_file(str(path), file_offset=8192)
    eng.close()


def test_group_k_larger_than_a_shard_pads_like_faiss():
    n, k = 40, 30  # 3 shards of 14/14/12 rows, k exceeds every shard
    db = synth.host_db(n, base_seed=87)
    q = synth.host_queries(2, 87, normalise=True)
    eng = engine.LocalEngine(n, devices=[0, 0, 0])
    eng.upload(0, db)
    eng.finalize()
    s, i = eng.search(q, k, mode=native.MODE_GEMV)
    _check(s, i, q, db, k)
    s, i = eng.search(q, 64, mode=native.MODE_GEMV)  # k > N: (-inf, -1) padding
    assert (i[:, n:] == -1).all() and np.isneginf(s[:, n:]).all() and (np.sort(i[:, :n], axis=1) == np.arange(n)).all()
    eng.close()


def test_long_fallback_queue_is_completed_by_the_synchronous_calls():
    """More queries fail their certificate than the exact-scan passes enqueued behind a tensor-core search cover
    (FCS_ASYNC_FALLBACK_QUERIES): fcs_search and the group finish the queue; the asynchronous call reports it."""
    n, k = 20000, 10
    rng = np.random.Generator(np.random.PCG64(3))
    db = synth.host_db(n, base_seed=51)
    dup = db[123][None, :] + 2e-4 * rng.standard_normal((6000, 128)).astype(np.float32)
    db[5000:11000] = dup / np.linalg.norm(dup, axis=1, keepdims=True)
    nbad = native.ASYNC_FALLBACK_QUERIES + 13
    q = np.ascontiguousarray(np.stack([db[5000 + 37 * j] for j in range(nbad)] + [synth.host_queries(1, 52 + j, normalise=True)[0] for j in range(19)]))
    h = native.Database(n, keep_bf16=True)
    h.upload(0, db)
    h.finalize()
    s, i = h.search(q, k, mode=native.MODE_TC)
    assert h.timing().last_tc_fallbacks >= nbad
    _check(s, i, q, db, k)
    # asynchronous call: finish reports the queue and completes it in the caller's buffers
    dev = torch.device("cuda:0")
    qd = torch.from_numpy(q).to(dev)
    sc = torch.empty((q.shape[0], k), dtype=torch.float32, device=dev)
    ids = torch.empty((q.shape[0], k), dtype=torch.int64, device=dev)
    st = torch.cuda.Stream(dev)
    h.search_device(qd.data_ptr(), q.shape[0], k, sc.data_ptr(), ids.data_ptr(), mode=native.MODE_TC, stream=st.cuda_stream)
    assert h.search_finish(st.cuda_stream) >= nbad
    _check(sc.cpu().numpy(), ids.cpu().numpy(), q, db, k)
    h.close()
    eng = engine.LocalEngine(n, devices=[0, 0], keep_bf16=True)  # the duplicate cluster straddles the shard boundary
    eng.upload(0, db)
    eng.finalize()
    s, i = eng.search(q, k, mode=native.MODE_TC)
    assert eng.group.last_fallbacks() >= nbad
    _check(s, i, q, db, k)
    eng.close()
