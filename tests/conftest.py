import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no GPU in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture
def tiny_faiss_db(tmp_path):
    """A database in the reference's .json layout, 50 domains."""
    n = 50
    rng = np.random.default_rng(2)
    from merizo_search_b200 import synth

    emb = synth.host_db(n, base_seed=9)
    emb.tofile(tmp_path / "t_raw_128d_norm.db")
    names = [f"AF-P{i:05d}-F1-model_v4_TED{i % 3 + 1:02d}" for i in range(n)]
    with open(tmp_path / "t_raw_128d.index_names", "wb") as fh:
        for nm in names:
            fh.write(nm.ljust(32).encode() + b"\n")
    seqs = ["".join(rng.choice(list("ACDEFGHIKLMNPQRSTVWY"), size=int(rng.integers(25, 60)))) for _ in range(n)]
    coords = [rng.standard_normal((len(s), 3)).astype(np.float32) for s in seqs]
    metas = [json.dumps({"cath": f"1.10.{i}.1"}) for i in range(n)]

    def pack(chunks, stem):
        off, idx = 0, []
        with open(tmp_path / f"t_{stem}.db", "wb") as fh:
            for c in chunks:
                fh.write(c)
                idx.append((off, off + len(c)))
                off += len(c)
        np.asarray(idx, dtype=np.int64).tofile(tmp_path / f"t_{stem}.index")

    pack([s.encode("ascii") for s in seqs], "seq")
    pack([c.tobytes() for c in coords], "ca")
    pack([m.encode("ascii") for m in metas], "metadata")
    info = {"dbfname_IP": "t_raw_128d_norm.db", "DB_SIZE": n, "DB_DIM": 128, "db_names_f": "t_raw_128d.index_names",
            "sif": "t_seq.index", "sdf": "t_seq.db", "cif": "t_ca.index", "cdf": "t_ca.db", "mif": "t_metadata.index",
            "mdf": "t_metadata.db"}
    with open(tmp_path / "t.json", "w") as fh:
        json.dump(info, fh)
    return tmp_path, emb, names, seqs, coords, metas


