"""Seeded synthetic Foldclass databases and queries.

Rows are i.i.d. N(0,1) fp32 [n,128], optionally L2-normalised, generated PER
262 144-ROW BLOCK from ``seed = base_seed + block_id`` so that any block can be
regenerated on the host for the oracle and shard-locally on a GPU rank (the
TED-scale matrix, 187 GB, never exists in one place).

Two generators with the same blocking:
  * ``host_block`` / ``host_db``   -- numpy (PCG64); reproducible anywhere.
  * ``device_block``               -- torch CUDA generator; used by bench.py for
    the multi-GB configs (values differ from the host generator; tests that
    need the oracle download the rows they check).
"""
from __future__ import annotations

import numpy as np

DIM = 128
BLOCK_ROWS = 262144  # the reference's default --search_batchsize (merizo.py:145)


def _normalise(x: np.ndarray) -> np.ndarray:
    n = np.sqrt((x.astype(np.float32) ** 2).sum(axis=1, keepdims=True, dtype=np.float32))
    return (x / np.maximum(n, np.float32(1e-12))).astype(np.float32)


def host_block(block_id: int, rows: int = BLOCK_ROWS, base_seed: int = 0, normalise: bool = True) -> np.ndarray:
    rng = np.random.Generator(np.random.PCG64(base_seed + block_id))
    x = rng.standard_normal((rows, DIM), dtype=np.float32)
    return _normalise(x) if normalise else x


def host_db(n_rows: int, base_seed: int = 0, normalise: bool = True) -> np.ndarray:
    out = np.empty((n_rows, DIM), dtype=np.float32)
    for b, r0 in enumerate(range(0, n_rows, BLOCK_ROWS)):
        r = min(BLOCK_ROWS, n_rows - r0)
        out[r0:r0 + r] = host_block(b, r, base_seed, normalise)
    return out


def host_queries(nq: int, batch_id: int = 0, normalise: bool = False, planted_from: np.ndarray | None = None,
                 planted_ids: np.ndarray | None = None, noise: float = 0.05) -> np.ndarray:
    """Queries from the same distribution (seed = 10**6 + batch_id).  If
    ``planted_from`` rows are given, query i is row ``planted_ids[i]`` plus small
    noise, so its top-1 is known."""
    rng = np.random.Generator(np.random.PCG64(10**6 + batch_id))
    q = rng.standard_normal((nq, DIM), dtype=np.float32)
    if planted_from is not None:
        ids = np.asarray(planted_ids)
        q[: len(ids)] = planted_from[ids] + noise * q[: len(ids)] / np.sqrt(np.float32(DIM))
    return _normalise(q) if normalise else q


def host_lengths(n_rows: int, seed: int = 7, lo: int = 25, hi: int = 683) -> np.ndarray:
    """Domain lengths in the range seen in the bundled TED slice (25..683)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    # log-normal-ish body like real domain lengths, clipped to the observed range
    x = np.exp(rng.normal(np.log(110.0), 0.5, size=n_rows))
    return np.clip(np.rint(x), lo, hi).astype(np.int32)


def device_block(block_id: int, rows: int, device, base_seed: int = 0, normalise: bool = True):
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(base_seed + block_id)
    x = torch.randn((rows, DIM), dtype=torch.float32, device=device, generator=g)
    if normalise:
        x = torch.nn.functional.normalize(x)
    return x


# ----------------------------------------------------------------------------------------------------------------
# Query-embedder inputs: seeded stand-in FoldClassNet(128) weights and C-alpha-like chains
# ----------------------------------------------------------------------------------------------------------------
EGNN_HID1 = 2 * (2 * DIM + 1)  # 514 = edge_input_dim * 2   (reference my_egnn_nocoords.py:14-21)
EGNN_MDIM = 2 * DIM            # 256 = m_dim                (reference nndef_fold_egnn_embed.py:46)
EGNN_MAX_LEN = 3000            # PositionalEncoder max_len  (reference nndef_fold_egnn_embed.py:13)
EGNN_N_LAYERS = 2
EGNN_LAYER_KEYS = ("edge_mlp.0.weight", "edge_mlp.0.bias", "edge_mlp.2.weight", "edge_mlp.2.bias",
                   "edge_gate.0.weight", "edge_gate.0.bias", "node_mlp.0.weight", "node_mlp.0.bias",
                   "node_mlp.2.weight", "node_mlp.2.bias")
EGNN_LAYER_SHAPES = {
    "edge_mlp.0.weight": (EGNN_HID1, 2 * DIM + 1), "edge_mlp.0.bias": (EGNN_HID1,),
    "edge_mlp.2.weight": (EGNN_MDIM, EGNN_HID1), "edge_mlp.2.bias": (EGNN_MDIM,),
    "edge_gate.0.weight": (1, EGNN_MDIM), "edge_gate.0.bias": (1,),
    "node_mlp.0.weight": (2 * DIM, DIM + EGNN_MDIM), "node_mlp.0.bias": (2 * DIM,),
    "node_mlp.2.weight": (DIM, 2 * DIM), "node_mlp.2.bias": (DIM,),
}


def positional_table(max_len: int = EGNN_MAX_LEN, d_model: int = DIM) -> np.ndarray:
    """nndef_fold_egnn_embed.py:15-20, evaluated in fp32 like torch does."""
    from math import log

    import torch  # torch's exp/sin/cos in fp32 are the reference arithmetic; numpy's differ in the last ulp

    pe = torch.zeros(max_len, d_model)
    position = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2).float() * (-log(10000.0) / d_model))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe.numpy()


def synthetic_state_dict(seed: int, dist_weight_scale: float = 0.002, message_weight_scale: float = 0.02) -> dict:
    """Seeded stand-in for FINAL_foldclass_model.pt (a missing large blob): weights ~ N(0, 1/fan_in) so that
    activations are O(1) through both layers (the module's own init, std 1e-3, would make every layer a
    near no-op and test nothing).  The column that multiplies dist^2 (values up to ~1e4 A^2) and the node-MLP
    columns that read the summed messages (a sum over up to hundreds of neighbours) are scaled down."""
    rng = np.random.Generator(np.random.PCG64(seed))
    sd = {"posenc_as.pe": positional_table()[None]}
    for layer in range(EGNN_N_LAYERS):
        for key in EGNN_LAYER_KEYS:
            shape = EGNN_LAYER_SHAPES[key]
            if key.endswith("weight"):
                w = rng.standard_normal(shape, dtype=np.float32) / np.float32(np.sqrt(shape[1]))
                if key == "edge_mlp.0.weight":
                    w[:, 2 * DIM] *= np.float32(dist_weight_scale)
                if key == "node_mlp.0.weight":
                    w[:, DIM:] *= np.float32(message_weight_scale)
            else:
                w = rng.standard_normal(shape, dtype=np.float32) * np.float32(0.1)
            sd[f"encode_ca_egnn.{layer}.{key}"] = w.astype(np.float32)
    return sd


def synthetic_chain(length: int, seed: int) -> np.ndarray:
    """A CA trace-like random walk: 3.8 A steps with persistence, fp32 [L,3]."""
    rng = np.random.Generator(np.random.PCG64(seed))
    d = rng.standard_normal((length, 3))
    for i in range(1, length):
        d[i] = 0.6 * d[i - 1] + 0.8 * d[i]
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return np.cumsum(3.8 * d, axis=0).astype(np.float32)




def synthetic_chains(lengths, seed: int = 0):
    """Vectorised batch of C-alpha-like chains (3.8 A steps with persistence), one per entry of `lengths`."""
    rng = np.random.Generator(np.random.PCG64(seed))
    out = []
    for L in lengths:
        d = rng.standard_normal((int(L), 3))
        # AR(1) filter along the chain without a Python loop per residue
        from scipy.signal import lfilter

        d = lfilter([0.8], [1.0, -0.6], d, axis=0)
        d /= np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-9)
        out.append(np.cumsum(3.8 * d, axis=0).astype(np.float32))
    return out
