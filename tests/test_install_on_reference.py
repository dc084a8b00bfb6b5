"""CPU, build container only (skipped where /root/reference is absent, e.g. on the GPU box): install() splices the
B200 path into the REAL reference module, and the reference's own driver then resolves the replaced callables."""
import importlib
import inspect
import os
import sys

import pytest

REF = "/root/reference/merizo_search"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")


@pytest.fixture()
def ref_module():
    sys.path.insert(0, REF)
    try:
        for name in [m for m in sys.modules if m.startswith("programs")]:
            del sys.modules[name]
        yield importlib.import_module("programs.Foldclass.dbsearch")
    finally:
        sys.path.remove(REF)
        for name in [m for m in sys.modules if m.startswith("programs")]:
            del sys.modules[name]


def test_install_replaces_the_hot_path_callables(ref_module):
    from merizo_search_b200 import dbsearch as b200
    from merizo_search_b200 import faiss_driver

    originals = {n: getattr(ref_module, n) for n in ("read_database", "search_query_against_db", "dbsearch_faiss", "network_setup")}
    b200.install(ref_module)
    assert ref_module.read_database is b200.read_database
    assert ref_module.search_query_against_db is b200.search_query_against_db
    assert ref_module.dbsearch_faiss is faiss_driver.dbsearch_faiss
    assert ref_module.network_setup.__wrapped__ is originals["network_setup"]
    # the reference's own drivers look these names up in the module namespace at call time
    src = inspect.getsource(ref_module.run_dbsearch)
    assert "read_database(" in src and "dbsearch_faiss(" in src and "network_setup(" in src
    assert "search_query_against_db(" in inspect.getsource(ref_module.dbsearch)


def test_replacements_keep_the_reference_signatures(ref_module):
    from merizo_search_b200 import dbsearch as b200
    from merizo_search_b200 import faiss_driver

    def params(f):
        return list(inspect.signature(f).parameters)

    assert params(b200.read_database) == params(ref_module.read_database)
    assert params(b200.search_query_against_db) == params(ref_module.search_query_against_db)
    assert params(faiss_driver.dbsearch_faiss) == params(ref_module.dbsearch_faiss)
    ref_defaults = {k: v.default for k, v in inspect.signature(ref_module.dbsearch_faiss).parameters.items()}
    our_defaults = {k: v.default for k, v in inspect.signature(faiss_driver.dbsearch_faiss).parameters.items()}
    assert our_defaults == ref_defaults


def test_embedder_accepts_the_reference_networks_state_dict(ref_module):
    """The keys and shapes FoldClassEmbedder reads are the ones the reference module owns (no checkpoint needed)."""
    from merizo_search_b200 import embed as b200_embed
    from merizo_search_b200 import native

    net = ref_module.FoldClassNet(128).eval()
    sd = net.state_dict()
    layers = b200_embed.layers_from_state_dict(sd)
    assert len(layers) == 2
    for layer in layers:
        for name, (_suffix, shape) in native.EGNN_KEYS.items():
            assert layer[name].shape == shape
    assert b200_embed.positional_table_from_state_dict(sd).shape == (3000, 128)
