"""GPU: the batched CUDA embedder (fcs_embed*, csrc/fcs_embed.cu) through the C ABI against
  * golden vectors made by the reference's own FoldClassNet (tests/golden/make_golden_embed.py), and
  * the numpy oracle on seeded ragged batches.
Parity rule (oracle.foldclass_embed_oracle.embedding_close): every component within 5e-5 of the largest
component of the reference embedding (fp32 network, different summation order), cosine >= 1 - 1e-6."""
import numpy as np
import pytest
import torch

from merizo_search_b200 import embed as b200_embed
from merizo_search_b200 import native
from oracle import foldclass_embed_oracle as emb
from golden_util import embed_golden as load_golden

pytestmark = pytest.mark.gpu
RTOL = 5e-5


@pytest.fixture(scope="module")
def golden():
    return load_golden()


@pytest.fixture(scope="module", params=["tc3", "tc2", "tc", "fp32"])
def embedder(golden, request):
    """All edge kernels behind the same ABI: the three tcgen05 kernels (bf16 hi/lo split; tc3 is the default) and the fp32
    FMA-pipe kernel."""
    _, sd, _ = golden
    e = b200_embed.FoldClassEmbedder(sd, device=0)
    assert e._emb.timing().last_launches == 0
    e._emb.set_mode({"tc3": native.EMBED_MODE_TC3, "tc2": native.EMBED_MODE_TC2, "tc": native.EMBED_MODE_TC,
                     "fp32": native.EMBED_MODE_FP32}[request.param])
    yield e
    e.close()


def _rel(got, want):
    return float(np.abs(got - want).max() / np.abs(want).max())


def test_layer_outputs_match_reference(golden, embedder):
    """Localises a failure: messages and node features of each EGNN layer for the first golden structure."""
    z, sd, structures = golden
    c = structures[0]
    f0 = sd["posenc_as.pe"][0, :c.shape[0]]
    want_l0, want_m0 = emb.egnn_layer(f0, c, sd, 0, factored=True, return_messages=True)
    want_l1, want_m1 = emb.egnn_layer(want_l0, c, sd, 1, factored=True, return_messages=True)
    got_l0, got_m0 = embedder._emb.debug_layer(c, 0)
    got_l1, got_m1 = embedder._emb.debug_layer(c, 1)
    errs = dict(m0=_rel(got_m0, want_m0), l0=_rel(got_l0, z["s0_layer0"]), m1=_rel(got_m1, want_m1), l1=_rel(got_l1, z["s0_layer1"]))
    print("relative errors per stage:", errs)
    if errs["m0"] > RTOL:  # localise: which (residue, channel) entries are off
        d = np.abs(got_m0 - want_m0) / np.abs(want_m0).max()
        bad = np.argwhere(d > RTOL)
        print("bad entries:", len(bad), "of", d.size, "rows:", sorted(set(bad[:, 0].tolist()))[:16], "cols:", sorted(set(bad[:, 1].tolist()))[:16])
    assert all(v < RTOL for v in errs.values()), errs


def test_golden_embeddings(golden, embedder):
    z, _, structures = golden
    got = embedder.embed_structures(structures)
    assert got.shape == (len(structures), 128) and got.dtype == np.float32
    bad = emb.embedding_close(got, z["embeddings"], rtol=RTOL)
    assert bad == [], bad
    t = embedder.timing()
    assert t.last_structures == len(structures) and t.last_residues == int(z["offsets"][-1])
    assert t.last_launches == 2 + 3 * 2 and t.last_edge_ms > 0


def test_one_structure_at_a_time_equals_batch(golden, embedder):
    _, _, structures = golden
    batch = embedder.embed_structures(structures[:8])
    for i in range(8):
        one = embedder.embed_structures([structures[i]])
        assert np.array_equal(one[0], batch[i]), f"structure {i}: result depends on the batch composition"


def test_ragged_batch_against_oracle(golden, embedder):
    _, sd, _ = golden
    rng = np.random.default_rng(11)
    lens = [int(x) for x in rng.integers(1, 200, size=24)] + [1, 8, 16, 24, 3000 // 10]
    structures = [emb.synthetic_chain(L, seed=500 + i) for i, L in enumerate(lens)]
    got = embedder.embed_structures(structures)
    want = emb.forward_batch(structures, sd, factored=True)
    bad = emb.embedding_close(got, want, rtol=RTOL)
    assert bad == [], bad


def test_module_call_contract(golden, embedder):
    """network(query_input): [1,L,3] tensor in, [1,128] tensor out, on the input's device (dbsearch.py:97-98)."""
    z, _, structures = golden
    x = torch.from_numpy(structures[0]).unsqueeze(0)
    y = embedder(x)
    assert isinstance(y, torch.Tensor) and tuple(y.shape) == (1, 128) and y.device == x.device
    assert emb.embedding_close(y.numpy(), z["embeddings"][:1], rtol=RTOL) == []
    yc = embedder(x.cuda())
    assert yc.is_cuda and np.array_equal(yc.cpu().numpy(), y.numpy())
    assert embedder.eval() is embedder and embedder.to("cuda") is embedder
    q = torch.zeros(3, 128)
    q[1, :] = embedder(x)  # the reference's query_embeddings[i,:] = network(query_input), dbsearch.py:301
    assert np.array_equal(q[1].numpy(), y[0].numpy())


def test_device_output_feeds_the_search(golden, embedder):
    """Embeddings stay in HBM and go straight into fcs_search_device: same hits as the host round trip."""
    z, _, structures = golden
    from merizo_search_b200 import synth

    q_dev = embedder.embed_structures_device(structures)
    host = embedder.embed_structures(structures)
    assert np.array_equal(q_dev.cpu().numpy(), host)
    n, k = 5000, 5
    db = synth.host_db(n, base_seed=3)
    h = native.Database(n)
    h.upload(0, db)
    h.finalize()
    sc = torch.empty((len(structures), k), dtype=torch.float32, device="cuda")
    ids = torch.empty((len(structures), k), dtype=torch.int64, device="cuda")
    st = torch.cuda.Stream()
    h.search_device(q_dev.data_ptr(), len(structures), k, sc.data_ptr(), ids.data_ptr(), qnorm=native.QNORM_L2,
                    mode=native.MODE_GEMV, stream=st.cuda_stream)
    st.synchronize()
    s2, i2 = h.search(host, k, qnorm=native.QNORM_L2, mode=native.MODE_GEMV)
    h.close()
    assert np.array_equal(ids.cpu().numpy(), i2) and np.array_equal(sc.cpu().numpy(), s2)


def test_many_structures_cross_the_pass_boundary(golden, embedder):
    """> 2^20 residues: the batch is split into passes; every copy of a structure must get the same embedding."""
    _, sd, _ = golden
    uniq = [emb.synthetic_chain(L, seed=800 + i) for i, L in enumerate([118, 119, 120, 121, 122, 123, 117])]
    want = emb.forward_batch(uniq, sd, factored=True)
    n = 9000
    structures = [uniq[i % len(uniq)] for i in range(n)]
    assert sum(c.shape[0] for c in structures) > (1 << 20)
    got = embedder.embed_structures(structures)
    for u in range(len(uniq)):
        rows = got[u::len(uniq)]
        assert np.array_equal(rows, np.broadcast_to(rows[0], rows.shape)), f"copies of structure {u} differ"
    assert emb.embedding_close(got[:len(uniq)], want, rtol=RTOL) == []
    t = embedder.timing()
    print(f"{n} structures, {t.last_residues} residues, {t.last_pairs} pairs: {t.last_ms:.1f} ms "
          f"(edge kernel {t.last_edge_ms:.1f} ms = {2 * 2 * 528 * 256 * t.last_pairs / (t.last_edge_ms * 1e-3) / 1e12:.1f} TFLOP/s fp32)")


def test_invalid_lengths_are_errors(embedder):
    with pytest.raises(native.FcsError):
        embedder.embed_structures([np.zeros((0, 3), np.float32)])
    with pytest.raises(native.FcsError):
        embedder.embed_structures([np.zeros((3001, 3), np.float32)])
    assert embedder.embed_structures([]).shape == (0, 128)
