"""Rebuild the inputs of the committed golden fixtures (tests/golden/*.npz; made by make_golden.py)."""
import hashlib
import os

import numpy as np

from merizo_search_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _cases(z):
    out = []
    for j in range(int(z["n_cases"])):
        out.append({key: z[f"c{j}_{key}"] for key in ("qi", "qlen", "mincov", "k", "scores", "ids")})
    for c in out:
        c["qi"], c["qlen"], c["k"], c["mincov"] = int(c["qi"]), int(c["qlen"]), int(c["k"]), float(c["mincov"])
    return out


def torch_flavour_n2048():
    z = np.load(os.path.join(GOLDEN, "torch_flavour_n2048.npz"))
    n = 2048
    db = synth.host_db(n, base_seed=int(z["db_seed"]), normalise=False)
    scale = np.exp(np.random.Generator(np.random.PCG64(int(z["scale_seed"]))).normal(0, 1.5, size=(n, 1))).astype(np.float32)
    db = (db * scale).astype(np.float32)
    db[17] = 0.0
    assert _sha(db) == str(z["db_sha"]), "synthetic DB drifted from the one the golden vectors were made on"
    return db, z["lengths"].astype(np.int32), z["queries"].astype(np.float32), _cases(z)


def torch_flavour_n300_full():
    z = np.load(os.path.join(GOLDEN, "torch_flavour_n300_full.npz"))
    return z["db"].astype(np.float32), z["lengths"].astype(np.int32), z["queries"].astype(np.float32), _cases(z)


def config1():
    z = np.load(os.path.join(GOLDEN, "config1_m0_vs_cath_size.npz"))
    db = synth.host_db(14942, base_seed=int(z["db_seed"]), normalise=False)
    lens = synth.host_lengths(14942, seed=int(z["len_seed"]))
    assert _sha(db) == str(z["db_sha"]) and _sha(lens) == str(z["lengths_sha"]), "synthetic inputs drifted"
    return db, lens, z


def ip_flavour():
    z = np.load(os.path.join(GOLDEN, "ip_flavour_n66943.npz"))
    db = synth.host_db(66943, base_seed=int(z["db_seed"]), normalise=True)
    assert _sha(db) == str(z["db_sha"]), "synthetic DB drifted"
    return db, z


def embed_golden():
    """(npz, seeded stand-in state_dict, list of [L,3] structures) of embed_foldclassnet.npz (make_golden_embed.py)."""
    from oracle import foldclass_embed_oracle as emb

    z = np.load(os.path.join(GOLDEN, "embed_foldclassnet.npz"))
    sd = emb.synthetic_state_dict(int(z["weight_seed"]))
    h = hashlib.sha256()
    for key in sorted(sd):
        h.update(key.encode())
        h.update(np.ascontiguousarray(sd[key]).tobytes())
    assert h.hexdigest() == str(z["weights_sha"]), "synthetic weights drifted from the ones the golden vectors were made with"
    offsets = z["offsets"]
    structures = [z["coords"][offsets[i]:offsets[i + 1]] for i in range(len(offsets) - 1)]
    return z, sd, structures
