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


def test_install_rebinds_modules_that_star_imported_the_originals(ref_module):
    """dbsearch_fulllength.py:29 does `from .dbsearch import *` and merizo.py:15 imports it at start-up, i.e. BEFORE
    install() can run: multi_domain_search (dbsearch_fulllength.py:303) must still get the resident database."""
    from merizo_search_b200 import dbsearch as b200
    from merizo_search_b200 import faiss_driver

    full = importlib.import_module("programs.Foldclass.dbsearch_fulllength")  # imported first, like merizo.py does
    assert full.read_database is ref_module.read_database
    b200.install(ref_module)
    assert full.read_database is b200.read_database
    assert full.search_query_against_db is b200.search_query_against_db
    assert full.dbsearch_faiss is faiss_driver.dbsearch_faiss
    assert full.network_setup is ref_module.network_setup and hasattr(full.network_setup, "__wrapped__")
    assert "read_database(" in inspect.getsource(full.multi_domain_search)


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


# ------------------------------------------------------------------------------------------------------------------
# Driver-level parity: the reference's OWN dbsearch() run twice on the same .pt database -- untouched, and with
# install() applied (the CUDA engine replaced by the CPU oracle, since this container has no GPU).  Everything around
# the replaced callables (query embedding, index lookups, metadata retrieval, result dicts) is the reference's code.
# ------------------------------------------------------------------------------------------------------------------
class _OraclePtEngine:
    """engine.LocalEngine look-alike for the .pt flavour: raw rows + lengths in, cosine * mask -> topk by the oracle."""

    def __init__(self, n_rows, devices=None, normalise_rows=False, keep_bf16=False, has_lengths=False):
        import numpy as np

        assert normalise_rows and has_lengths, "the .pt flavour uploads raw rows and lengths"
        self.n_rows = n_rows
        self.rows = np.zeros((n_rows, 128), np.float32)
        self.lens = np.zeros(n_rows, np.int32)

    def upload(self, row0, rows, lengths=None):
        self.rows[row0:row0 + len(rows)] = rows
        self.lens[row0:row0 + len(rows)] = lengths

    def finalize(self):
        pass

    def close(self):
        pass

    def search(self, q, k, qlen=None, mincov=0.0, qnorm=None, mode=None, kprime=0):
        import numpy as np
        import torch

        from merizo_search_b200 import native
        from oracle import foldclass_oracle as orc

        assert qnorm == native.QNORM_COSINE and q.shape[0] == 1
        s, i, _ = orc.search_torch_flavour(torch.from_numpy(self.rows), torch.from_numpy(self.lens.astype(np.float32)),
                                           torch.from_numpy(q[0]), int(qlen[0]), float(mincov), int(k))
        return s.numpy()[None], i.numpy()[None]


def test_reference_dbsearch_driver_gives_the_same_hits_with_the_spliced_path(ref_module, tmp_path, monkeypatch):
    import json
    import pickle

    import numpy as np
    import torch

    from merizo_search_b200 import dbsearch as b200
    from merizo_search_b200 import synth

    torch.manual_seed(3)
    net = ref_module.FoldClassNet(128).eval()
    with torch.no_grad():  # give the random-init network O(1) weights so that embeddings differ between structures
        for prm in net.parameters():
            prm.mul_(300.0)
    lens = [30, 45, 25, 60, 33, 28, 51, 40, 37, 26, 48, 55]
    chains = synth.synthetic_chains(lens, seed=17)
    seqs = ["A" * L for L in lens]
    base = str(tmp_path / "mini")
    with torch.no_grad():
        rows = torch.cat([net(torch.from_numpy(c).unsqueeze(0)) for c in chains], dim=0)
    torch.save(rows, base + ".pt")
    with open(base + ".index", "wb") as fh:
        pickle.dump([(f"/x/dom{i:03d}.pdb", c, s) for i, (c, s) in enumerate(zip(chains, seqs))], fh)
    metas = [json.dumps({"cath": f"1.10.{i}.1"}) for i in range(len(lens))]
    off, idx = 0, []
    with open(base + ".metadata", "wb") as fh:
        for m in metas:
            fh.write(m.encode("ascii"))
            idx.append((off, off + len(m)))
            off += len(m)
    np.asarray(idx, dtype=np.int64).tofile(base + ".metadata.index")
    query = {"name": "/q/query.pdb", "coords": (chains[3] + 0.05).astype(np.float32), "seq": "A" * lens[3]}

    def run(read_device="cpu"):
        target = ref_module.read_database(base, torch.device(read_device))
        return ref_module.dbsearch(dict(query), target, str(tmp_path), net, 5, 0.7, 0.5, 0.5, False, torch.device("cpu"),
                                   inputs_are_ca=True, skip_tmalign=True)

    want, _ = run()                                   # the untouched reference
    monkeypatch.setattr(b200, "LocalEngine", _OraclePtEngine)
    b200._RESIDENT.clear()
    b200.install(ref_module)
    with pytest.raises(b200.native.FcsError):          # the spliced path has no CPU implementation and says so
        run()
    got, _ = run("cuda")                              # the same driver, hot path spliced (engine stubbed: no GPU touched)
    b200._RESIDENT.clear()
    assert len(want) >= 2 and list(got.keys()) == list(want.keys())
    for key in want:
        w, g = want[key], got[key]
        assert set(g) == set(w)
        for field in ("query", "target", "q_len", "t_len", "metadata", "tmalign_output", "dom_str"):
            assert g[field] == w[field], field
        assert int(g["dbindex"]) == int(w["dbindex"])
        assert abs(float(g["score"]) - float(w["score"])) <= 1e-6
        assert "{:.4f}".format(g["score"]) == "{:.4f}".format(w["score"])  # what the TSV writer prints


def test_record_readers_on_the_bundled_ted_slice_fixtures(tmp_path):
    """The real on-disk layout (examples/database/ted100_9606_small: names + offset indices are bundled, the .db payloads
    are missing large blobs): 33-byte name records, int64 [N,2] (start,end) tables, 12 bytes of coordinates per residue."""
    import json

    import numpy as np

    from merizo_search_b200 import faiss_driver

    src = "/root/reference/examples/database/ted100_9606_small"
    if not os.path.isdir(src):
        pytest.skip("bundled TED slice not present")
    info = json.load(open(os.path.join(src, "ted100_9606_small.json")))
    n = int(info["DB_SIZE"])
    assert n == 66943 and int(info["DB_DIM"]) == 128
    seq_idx = np.fromfile(os.path.join(src, info["sif"]), dtype=np.int64).reshape(-1, 2)
    ca_idx = np.fromfile(os.path.join(src, info["cif"]), dtype=np.int64).reshape(-1, 2)
    assert seq_idx.shape == (n, 2) and ca_idx.shape == (n, 2)
    lens = seq_idx[:, 1] - seq_idx[:, 0]
    assert lens.min() == 25 and lens.max() == 683
    assert np.array_equal(ca_idx[:, 1] - ca_idx[:, 0], 12 * lens)      # float32 [L,3] per domain
    assert np.array_equal(seq_idx[1:, 0], seq_idx[:-1, 1])             # records are contiguous
    # the readers need the payload files to exist: stand-ins of the right size, real names + indices
    for key in ("db_names_f", "sif", "cif", "mif"):
        os.symlink(os.path.join(src, info[key]), tmp_path / info[key])
    rng = np.random.default_rng(0)
    seq_payload = rng.choice(np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY", dtype=np.uint8), size=int(seq_idx[-1, 1]))
    seq_payload.tofile(tmp_path / info["sdf"])
    rec = faiss_driver.RecordFiles(str(tmp_path), info)
    ids = np.array([0, 1, 12345, n - 1])
    names = rec.names(ids)
    assert all(nm.startswith("AF-") and "TED" in nm and " " not in nm for nm in names), names
    seqs = rec.sequences(ids)
    assert [len(s) for s in seqs] == lens[ids].tolist()
    assert seqs[2] == bytes(seq_payload[seq_idx[12345, 0]:seq_idx[12345, 1]]).decode("ascii")
