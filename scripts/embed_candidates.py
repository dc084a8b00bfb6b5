"""GPU: parity + throughput of the embedder's edge-kernel modes (1 = tcgen05 kernel on main; 2 / 3 = the r2-prep candidates:
dedicated epilogue warps + double-buffered accumulators, with 8 / 16 generator warps).  Run with the candidate library:
    FCS_LIB_VARIANT=emb2 FCS_NO_REBUILD=1 python scripts/embed_candidates.py
Each mode runs in a child process with a timeout (an untested kernel may hang)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def child(mode):
    import numpy as np
    import torch

    from golden_util import embed_golden
    from merizo_search_b200 import embed as b200_embed
    from merizo_search_b200 import native, synth
    from oracle import foldclass_embed_oracle as eorc

    z, sd, structures = embed_golden()
    e = b200_embed.FoldClassEmbedder(sd, device=0)
    e._emb.set_mode(mode)
    got = e.embed_structures(structures)
    bad = eorc.embedding_close(got, z["embeddings"], rtol=5e-5)  # golden vectors made by the reference's FoldClassNet
    lens = synth.host_lengths(600, seed=11)
    chains = synth.synthetic_chains(lens, seed=7)
    got2 = e.embed_structures(chains)
    e._emb.set_mode(native.EMBED_MODE_FP32)
    ref2 = e.embed_structures(chains)  # the fp32 kernel as an independent path on a bigger ragged batch
    e._emb.set_mode(mode)
    bad2 = eorc.embedding_close(got2, ref2, rtol=5e-5)
    # throughput: 2048 structures like bench.py --workload embed
    lens = synth.host_lengths(2048, seed=21)
    chains = synth.synthetic_chains(lens, seed=3)
    coords, offsets = native.Embedder._pack(chains)
    out = torch.empty((2048, 128), dtype=torch.float32, device="cuda:0")
    for _ in range(3):
        e._emb.embed_packed_to_device(coords, offsets, out.data_ptr())
    ms = []
    for _ in range(3):
        e._emb.embed_packed_to_device(coords, offsets, out.data_ptr())
        ms.append(e.timing().last_ms)
    print(json.dumps({"mode": mode, "golden_violations": bad[:3], "batch_violations": bad2[:3], "ms_per_2048": float(np.mean(ms)),
                      "structures_per_s": 2048 / (float(np.mean(ms)) * 1e-3)}))
    e.close()


if __name__ == "__main__":
    if len(sys.argv) > 1:
        child(int(sys.argv[1]))
    else:
        for mode in (1, 2, 3):
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), str(mode)], capture_output=True, text=True, timeout=100)
                print(r.stdout.strip().splitlines()[-1] if r.stdout.strip() else f"mode {mode}: no output; stderr: {r.stderr[-400:]}")
            except subprocess.TimeoutExpired:
                print(f"mode {mode}: TIMEOUT (hang)")
