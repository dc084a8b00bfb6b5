"""GPU: BASELINE.json's full sizes.  The oracle cannot finish these in seconds, so the checks are size-independent
properties: planted near-duplicates must come back first, reported scores must equal an independent fp32
recomputation of the reported rows, the two independent kernel paths (fp32 scan vs tcgen05 + rescore) must agree
id-for-id, and on a random sample of queries the result must equal a brute-force torch top-k on the GPU."""
import numpy as np
import pytest
import torch

from merizo_search_b200 import native, synth

pytestmark = pytest.mark.gpu
BLK = 1 << 20


def _build(n, dev, keep_bf16, planted):
    """Device-generated unit-norm DB (block seeds as in bench.py); returns handle + the planted rows (fp32, host)."""
    h = native.Database(n, keep_bf16=keep_bf16)
    keep = {}
    for b0 in range(0, n, BLK):
        nb = min(BLK, n - b0)
        x = synth.device_block(b0 // BLK, nb, dev, base_seed=2000)
        for r in planted:
            if b0 <= r < b0 + nb:
                keep[r] = x[r - b0].cpu().numpy().copy()
        h.upload_device(b0, nb, x.data_ptr())
        del x
    h.finalize()
    return h, keep


def _brute_force(n, dev, q, k):
    """Exact fp32 top-k by regenerating the DB block by block with torch (an independent code path)."""
    best_s = torch.full((q.shape[0], k), -float("inf"), device=dev)
    best_i = torch.full((q.shape[0], k), -1, dtype=torch.int64, device=dev)
    for b0 in range(0, n, BLK):
        nb = min(BLK, n - b0)
        x = synth.device_block(b0 // BLK, nb, dev, base_seed=2000)
        s = q @ x.T
        ts, ti = torch.topk(s, min(k, nb), dim=1)
        cs, ci = torch.cat([best_s, ts], 1), torch.cat([best_i, ti + b0], 1)
        best_s, pos = torch.topk(cs, k, dim=1)
        best_i = torch.gather(ci, 1, pos)
    return best_s.cpu().numpy(), best_i.cpu().numpy()


def _agree(s, i, ws, wi, tol=1e-5):
    assert np.abs(s - ws).max() <= tol, f"score diff {np.abs(s - ws).max():.2e}"
    mism = i != wi
    if mism.any():  # only exact-tie permutations are acceptable
        assert (np.abs(s[mism] - ws[mism]) <= tol).all()
        assert mism.mean() < 0.01, "too many id mismatches for ties"


def test_config2_cath_scale_single_query_k10():
    dev = torch.device("cuda:0")
    n, k = 500_000, 10
    planted = [7, 250_000, n - 1]
    h, rows = _build(n, dev, False, planted)
    rng = np.random.default_rng(0)
    for r in planted:
        q = rows[r] + 0.02 * rng.standard_normal(128).astype(np.float32) / np.sqrt(128)
        s, i = h.search(q, k, qnorm=native.QNORM_L2, mode=native.MODE_GEMV)
        assert i[0, 0] == r and s[0, 0] > 0.99
        assert (np.diff(s[0]) <= 0).all() and len(set(i[0])) == k
    qs = torch.nn.functional.normalize(torch.randn(4, 128, device=dev, generator=torch.Generator(dev).manual_seed(1)))
    s, i = h.search(qs.cpu().numpy(), k, mode=native.MODE_GEMV)
    ws, wi = _brute_force(n, dev, qs, k)
    _agree(s, i, ws, wi)
    h.close()


def test_config3_10m_rows_4096_queries_k100_tensor_core_path():
    dev = torch.device("cuda:0")
    n, nq, k = 10_000_000, 4096, 100
    planted = [3, 5_000_000, n - 1]
    h, rows = _build(n, dev, True, planted)
    g = torch.Generator(dev).manual_seed(2)
    q = torch.nn.functional.normalize(torch.randn(nq, 128, device=dev, generator=g)).cpu().numpy()
    for j, r in enumerate(planted):
        q[j] = rows[r]  # exact duplicates of database rows
    s, i = h.search(q, k, mode=native.MODE_TC)
    t = h.timing()
    assert t.last_mode == native.MODE_TC and t.last_tc_fallbacks <= 8
    for j, r in enumerate(planted):
        assert i[j, 0] == r and abs(s[j, 0] - 1.0) < 1e-5
    assert (np.diff(s, axis=1) <= 0).all() and (i >= 0).all() and (i < n).all()
    assert all(len(set(row)) == k for row in i[:: 64])
    # the independent exact fp32 scan on a sample of the queries must agree id-for-id
    sample = np.r_[0:3, 1000:1013]
    s2, i2 = h.search(q[sample], k, mode=native.MODE_GEMV)
    _agree(s[sample], i[sample], s2, i2, tol=2e-6)
    # and so must a brute-force torch top-k
    ws, wi = _brute_force(n, dev, torch.from_numpy(q[sample[:6]]).to(dev), k)
    _agree(s[sample[:6]], i[sample[:6]], ws, wi)
    h.close()


def test_config4_ted_slice_45m_rows_per_gpu():
    dev = torch.device("cuda:0")
    n, k = 45_625_000, 10
    free, _ = torch.cuda.mem_get_info(dev)
    if free < 45e9:
        pytest.skip("needs ~37 GB of free HBM")
    planted = [11, 30_000_000, n - 1]
    h, rows = _build(n, dev, True, planted)
    # batch 1 (exact scan)
    for r in planted:
        s, i = h.search(rows[r], k, mode=native.MODE_GEMV)
        assert i[0, 0] == r and abs(s[0, 0] - 1.0) < 1e-5
    # batch 1024 (tensor-core path) against the exact scan on a sample
    g = torch.Generator(dev).manual_seed(3)
    q = torch.nn.functional.normalize(torch.randn(1024, 128, device=dev, generator=g)).cpu().numpy()
    q[5] = rows[planted[1]]
    s, i = h.search(q, k, mode=native.MODE_TC)
    assert i[5, 0] == planted[1]
    sample = np.arange(0, 16)
    s2, i2 = h.search(q[sample], k, mode=native.MODE_GEMV)
    _agree(s[sample], i[sample], s2, i2, tol=2e-6)
    assert h.timing().last_mode == native.MODE_GEMV
    h.close()


def test_config5_style_proteome_batch_65536_queries_k50():
    """BASELINE configs[4] shape in the batch dimension (65 536 queries, k=50) against a 2 M-row shard: 128 query
    groups, 2 GB of candidate buffers -- the plumbing the 8-GPU proteome-wide search needs on every rank."""
    dev = torch.device("cuda:0")
    n, nq, k = 2_000_000, 65_536, 50
    h, rows = _build(n, dev, True, [17])
    g = torch.Generator(dev).manual_seed(5)
    q = torch.nn.functional.normalize(torch.randn(nq, 128, device=dev, generator=g)).cpu().numpy()
    q[40000] = rows[17]
    s, i = h.search(q, k, mode=native.MODE_TC)
    assert h.timing().last_mode == native.MODE_TC and h.timing().last_tc_fallbacks <= 64
    assert i[40000, 0] == 17 and abs(s[40000, 0] - 1.0) < 1e-5
    assert (np.diff(s, axis=1) <= 0).all() and (i >= 0).all() and (i < n).all()
    sample = np.array([0, 511, 512, 40000, 65535, 33333, 12345, 60000])
    s2, i2 = h.search(q[sample], k, mode=native.MODE_GEMV)
    _agree(s[sample], i[sample], s2, i2, tol=2e-6)
    h.close()
