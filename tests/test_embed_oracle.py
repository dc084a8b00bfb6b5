"""CPU: the embedder's oracle (numpy restatement of FoldClassNet.forward) against golden vectors produced by
the reference's own module (tests/golden/make_golden_embed.py), plus the host-side plumbing of the drop-in."""
import types

import numpy as np
import pytest

from merizo_search_b200 import embed as b200_embed
from merizo_search_b200 import native
from oracle import foldclass_embed_oracle as emb

from golden_util import embed_golden as load_golden  # noqa: E402


@pytest.mark.parametrize("factored", [False, True])
def test_oracle_matches_reference_network(factored):
    z, sd, structures = load_golden()
    # the long chains cost seconds in numpy: keep every tile-boundary length, skip only the 400-residue one
    keep = [i for i, c in enumerate(structures) if c.shape[0] <= 300]
    got = emb.forward_batch([structures[i] for i in keep], sd, factored=factored)
    want = z["embeddings"][keep]
    assert emb.embedding_close(got, want, rtol=5e-6) == []


def test_oracle_layer_outputs_match_reference():
    z, sd, structures = load_golden()
    c = structures[0]
    f = sd["posenc_as.pe"][0, :c.shape[0]]
    l0 = emb.egnn_layer(f, c, sd, 0, factored=True)
    l1 = emb.egnn_layer(l0, c, sd, 1, factored=True)
    np.testing.assert_allclose(l0, z["s0_layer0"], rtol=0, atol=5e-6 * np.abs(z["s0_layer0"]).max())
    np.testing.assert_allclose(l1, z["s0_layer1"], rtol=0, atol=5e-6 * np.abs(z["s0_layer1"]).max())
    np.testing.assert_allclose(l1.mean(0), z["embeddings"][0], rtol=0, atol=5e-6 * np.abs(z["embeddings"][0]).max())


def test_closeness_rule_rejects_small_errors():
    z, _, _ = load_golden()
    want = z["embeddings"]
    assert emb.embedding_close(want, want) == []
    off = want.copy()
    off[3, 17] += 1e-3 * np.abs(want[3]).max()
    assert len(emb.embedding_close(off, want)) == 1


def test_pack_layout():
    cs = [np.ones((3, 3), np.float32), np.zeros((1, 3), np.float32), 2 * np.ones((5, 3), np.float32)]
    coords, offsets = emb.pack(cs)
    assert offsets.tolist() == [0, 3, 4, 9] and coords.shape == (9, 3) and coords.dtype == np.float32
    c2, o2 = native.Embedder._pack(cs)
    assert np.array_equal(coords, c2) and np.array_equal(offsets, o2)


def test_state_dict_mapping_and_shapes():
    sd = emb.synthetic_state_dict(3)
    layers = b200_embed.layers_from_state_dict(sd)
    assert len(layers) == 2
    for layer in layers:
        assert set(layer) == set(native.EgnnWeights.FIELDS)
        for name, (_suffix, shape) in native.EGNN_KEYS.items():
            assert layer[name].shape == shape and layer[name].dtype == np.float32
    assert b200_embed.positional_table_from_state_dict(sd).shape == (3000, 128)
    with pytest.raises(KeyError):
        b200_embed.layers_from_state_dict({k: v for k, v in sd.items() if "edge_gate" not in k})


def test_embedder_needs_cuda_device():
    with pytest.raises(native.FcsError):
        b200_embed.FoldClassEmbedder(emb.synthetic_state_dict(3), device="cpu")
    import torch

    if not torch.cuda.is_available():
        with pytest.raises(native.FcsError):  # no GPU: the product path fails loudly, there is no CPU fallback
            b200_embed.FoldClassEmbedder(emb.synthetic_state_dict(3), device="cuda:0")


def test_network_setup_wrapper_only_touches_cuda(monkeypatch):
    made = {}

    class FakeEmbedder:
        @classmethod
        def from_network(cls, network, device):
            made["args"] = (network, device)
            return "EMBEDDER"

    monkeypatch.setattr(b200_embed, "FoldClassEmbedder", FakeEmbedder)
    ref = lambda threads, device: ("TORCH_NET", device)  # noqa: E731
    wrapped = b200_embed.wrap_network_setup(ref)
    assert wrapped(4, "cpu") == ("TORCH_NET", "cpu")
    assert wrapped(4, "cuda") == ("EMBEDDER", "cuda") and made["args"] == ("TORCH_NET", "cuda")
    assert wrapped.__wrapped__ is ref


def test_install_wraps_network_setup_once():
    from merizo_search_b200 import dbsearch as b200

    ref = types.SimpleNamespace(read_database=None, search_query_against_db=None, dbsearch_faiss=None,
                                network_setup=lambda threads, device: ("NET", device))
    b200.install(ref)
    first = ref.network_setup
    assert hasattr(first, "__wrapped__")
    b200.install(ref)
    assert ref.network_setup is first
