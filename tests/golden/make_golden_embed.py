#!/usr/bin/env python
"""Generate tests/golden/embed_foldclassnet.npz by running the REFERENCE's own FoldClassNet.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_embed.py

Imports ``FoldClassNet`` (nndef_fold_egnn_embed.py:34-62, my_egnn_nocoords.py:10-74) and ``read_pdb``
(Foldclass/utils.py:42) from /root/reference, loads the seeded stand-in weights of
``oracle.foldclass_embed_oracle.synthetic_state_dict`` (the trained FINAL_foldclass_model.pt is a missing
large blob) with ``load_state_dict(strict=True)``, and records the embeddings of the bundled example
structures and of seeded synthetic chains whose lengths straddle the kernels' tile sizes.  The structures'
coordinates are stored (a few KB); the weights are regenerated from the seed, with a sha256 guard.
"""
import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/merizo_search")

from programs.Foldclass.nndef_fold_egnn_embed import FoldClassNet  # noqa: E402
from programs.Foldclass.utils import read_pdb  # noqa: E402

from oracle import foldclass_embed_oracle as emb  # noqa: E402

WEIGHT_SEED = 2024
SYNTH_LENGTHS = [1, 2, 7, 8, 9, 15, 16, 17, 31, 33, 64, 65, 100, 127, 128, 129, 150, 257, 400]


def weights_sha(sd) -> str:
    h = hashlib.sha256()
    for key in sorted(sd):
        h.update(key.encode())
        h.update(np.ascontiguousarray(sd[key]).tobytes())
    return h.hexdigest()


def main():
    torch.manual_seed(0)
    torch.set_num_threads(1)
    sd = emb.synthetic_state_dict(WEIGHT_SEED)
    net = FoldClassNet(128).eval()
    missing = net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
    print("load_state_dict:", missing)

    structures, names = [], []
    for pdb, chain in [("M0.pdb", "A"), ("3w5h.pdb", "A")]:
        d = read_pdb(pdbfile=os.path.join("/root/reference/examples", pdb), pdb_chain=chain)
        structures.append(np.ascontiguousarray(d["coords"], dtype=np.float32))
        names.append(pdb)
    for i, L in enumerate(SYNTH_LENGTHS):
        structures.append(emb.synthetic_chain(L, seed=9000 + i))
        names.append(f"synthetic_L{L}")

    outs = []
    with torch.no_grad():
        for c in structures:
            outs.append(net(torch.from_numpy(c).unsqueeze(0)).numpy().reshape(-1).astype(np.float32))
    outs = np.stack(outs)
    # intermediates of one structure (after each EGNN layer), to localise a failing kernel
    with torch.no_grad():
        x = torch.from_numpy(structures[0]).unsqueeze(0)
        f0 = net.posenc_as(x)
        l0 = net.encode_ca_egnn[0]((f0, x, None))[0]
        l1 = net.encode_ca_egnn[1]((l0, x, None))[0]
    coords, offsets = emb.pack(structures)
    np.savez_compressed(os.path.join(HERE, "embed_foldclassnet.npz"), weight_seed=WEIGHT_SEED, weights_sha=weights_sha(sd),
                        coords=coords, offsets=offsets, names=np.array(names), embeddings=outs,
                        s0_layer0=l0.numpy()[0].astype(np.float32), s0_layer1=l1.numpy()[0].astype(np.float32))
    print("structures:", len(structures), "lengths:", np.diff(offsets).tolist())
    print("max |emb|:", np.abs(outs).max(), " mean |emb|:", np.abs(outs).mean())
    # how far the numpy restatement is from the reference (both orders), for the record
    for factored in (False, True):
        got = emb.forward_batch(structures, sd, factored=factored)
        rel = np.abs(got - outs).max(axis=1) / np.abs(outs).max(axis=1)
        print(f"oracle factored={factored}: max rel err {rel.max():.3e}; violations:", emb.embedding_close(got, outs))


if __name__ == "__main__":
    main()
