"""CPU oracle for the Foldclass query embedder -- TEST INFRASTRUCTURE ONLY.

Restates, in plain numpy fp32, the forward pass of the reference's ``FoldClassNet(128)`` in eval mode:

  nndef_fold_egnn_embed.py:11-30   PositionalEncoder (fixed sinusoidal table, ``learned=False``; the INPUT
                                   only supplies the length: forward() returns pe[:, :L, :])
  nndef_fold_egnn_embed.py:50-62   FoldClassNet.forward: feats = pe[:L]; two EGNN layers; mean over residues
  my_egnn_nocoords.py:44-74        EGNN.forward: edge_input = [feats_i, feats_j, dist*dist];
                                   m_ij = edge_mlp(edge_input); m_ij *= sigmoid(Linear(m_ij)); m_i = sum_j m_ij;
                                   node_out = node_mlp([feats, m_i]) + feats

Only ``tests/``, ``__graft_entry__.smoke()`` and bench.py's CPU legs may import this module; the product
package never does.  Pinned against the reference itself: ``tests/golden/make_golden_embed.py`` runs the
reference's ``FoldClassNet`` (imported from /root/reference in the build container) on seeded weights and
structures and commits its outputs; ``tests/test_embed_oracle.py`` checks this restatement against them.

Two evaluation orders are offered:
  * ``forward(..., factored=False)`` -- literally the reference's: materialise [L,L,257], two dense layers.
  * ``forward(..., factored=True)``  -- the algebraically identical split the CUDA kernels use:
        W1 @ [f_i, f_j, d2] + b1 = (W1[:, :128] f_i + b1) + (W1[:, 128:256] f_j) + d2 * W1[:, 256]
    (differs from the literal order by fp32 rounding only).
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np

WIDTH = 128
HID1 = 2 * (2 * WIDTH + 1)  # 514 = edge_input_dim * 2   (my_egnn_nocoords.py:18-21)
MDIM = 2 * WIDTH            # 256 = m_dim                (nndef_fold_egnn_embed.py:46)
MAX_LEN = 3000              # PositionalEncoder max_len  (nndef_fold_egnn_embed.py:13)
N_LAYERS = 2

LAYER_KEYS = ("edge_mlp.0.weight", "edge_mlp.0.bias", "edge_mlp.2.weight", "edge_mlp.2.bias",
              "edge_gate.0.weight", "edge_gate.0.bias", "node_mlp.0.weight", "node_mlp.0.bias",
              "node_mlp.2.weight", "node_mlp.2.bias")
LAYER_SHAPES = {
    "edge_mlp.0.weight": (HID1, 2 * WIDTH + 1), "edge_mlp.0.bias": (HID1,),
    "edge_mlp.2.weight": (MDIM, HID1), "edge_mlp.2.bias": (MDIM,),
    "edge_gate.0.weight": (1, MDIM), "edge_gate.0.bias": (1,),
    "node_mlp.0.weight": (2 * WIDTH, WIDTH + MDIM), "node_mlp.0.bias": (2 * WIDTH,),
    "node_mlp.2.weight": (WIDTH, 2 * WIDTH), "node_mlp.2.bias": (WIDTH,),
}


# Seeded stand-in weights / structures live in the package's synthetic-data module (bench.py uses them without
# touching the oracle); re-exported here because the golden fixtures and the tests address them through the oracle.
from merizo_search_b200.synth import positional_table, synthetic_chain, synthetic_state_dict  # noqa: E402,F401


def _silu(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):  # exp(-x) -> inf for very negative x: x / inf = -0, as in torch
        return (x / (np.float32(1) + np.exp(-x, dtype=np.float32))).astype(np.float32)


def _sigmoid(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        return (np.float32(1) / (np.float32(1) + np.exp(-x, dtype=np.float32))).astype(np.float32)


def egnn_layer(feats: np.ndarray, coords: np.ndarray, sd: Dict[str, np.ndarray], layer: int, factored: bool,
               return_messages: bool = False):
    """my_egnn_nocoords.py:44-74 for one structure: feats [L,128], coords [L,3] -> [L,128]."""
    p = f"encode_ca_egnn.{layer}."
    w1, b1 = sd[p + "edge_mlp.0.weight"], sd[p + "edge_mlp.0.bias"]
    w2, b2 = sd[p + "edge_mlp.2.weight"], sd[p + "edge_mlp.2.bias"]
    wg, bg = sd[p + "edge_gate.0.weight"], sd[p + "edge_gate.0.bias"]
    w3, b3 = sd[p + "node_mlp.0.weight"], sd[p + "node_mlp.0.bias"]
    w4, b4 = sd[p + "node_mlp.2.weight"], sd[p + "node_mlp.2.bias"]
    L = feats.shape[0]
    rel = coords[:, None, :] - coords[None, :, :]                       # :48
    dist = np.sqrt((rel * rel).sum(-1, dtype=np.float32), dtype=np.float32)  # :49 (linalg.norm)
    d2 = (dist * dist).astype(np.float32)                               # :58 (dist*dist, not the raw sum)
    if factored:
        pi = feats @ w1[:, :WIDTH].T + b1                               # [L,514]
        qj = feats @ w1[:, WIDTH:2 * WIDTH].T                           # [L,514]
        pre = pi[:, None, :] + qj[None, :, :] + d2[:, :, None] * w1[:, 2 * WIDTH][None, None, :]
    else:
        edge_input = np.concatenate([np.broadcast_to(feats[:, None, :], (L, L, WIDTH)),
                                     np.broadcast_to(feats[None, :, :], (L, L, WIDTH)), d2[:, :, None]], axis=-1)
        pre = edge_input @ w1.T + b1
    h1 = _silu(pre.astype(np.float32))                                  # edge_mlp[0..1]
    m = _silu((h1 @ w2.T + b2).astype(np.float32))                      # edge_mlp[2..3]
    gate = _sigmoid((m @ wg.T + bg).astype(np.float32))                 # :64 edge_gate
    m = m * gate
    m_i = m.sum(axis=1, dtype=np.float32)                               # :69 sum over j
    node_in = np.concatenate([feats, m_i], axis=-1)                     # :71
    n1 = _silu((node_in @ w3.T + b3).astype(np.float32))
    out = (n1 @ w4.T + b4 + feats).astype(np.float32)                   # :72 residual
    return (out, m_i) if return_messages else out


def forward(coords: np.ndarray, sd: Dict[str, np.ndarray], factored: bool = True) -> np.ndarray:
    """FoldClassNet.forward (nndef_fold_egnn_embed.py:50-62) for ONE structure: coords [L,3] -> [128]."""
    coords = np.ascontiguousarray(coords, dtype=np.float32).reshape(-1, 3)
    L = coords.shape[0]
    feats = sd["posenc_as.pe"].reshape(-1, WIDTH)[:L].astype(np.float32)
    for layer in range(N_LAYERS):
        feats = egnn_layer(feats, coords, sd, layer, factored)
    return feats.mean(axis=0, dtype=np.float32)


def forward_batch(structures: Sequence[np.ndarray], sd: Dict[str, np.ndarray], factored: bool = True) -> np.ndarray:
    return np.stack([forward(c, sd, factored) for c in structures]).astype(np.float32)


def pack(structures: Sequence[np.ndarray]):
    """Ragged batch -> (coords [sum L, 3] fp32, offsets int64 [n+1]) -- the layout the C ABI takes."""
    lens = [int(np.asarray(c).reshape(-1, 3).shape[0]) for c in structures]
    offsets = np.zeros(len(lens) + 1, dtype=np.int64)
    offsets[1:] = np.cumsum(lens)
    coords = np.concatenate([np.asarray(c, dtype=np.float32).reshape(-1, 3) for c in structures]) if lens else \
        np.zeros((0, 3), np.float32)
    return np.ascontiguousarray(coords, dtype=np.float32), offsets


def embedding_close(got: np.ndarray, want: np.ndarray, rtol: float = 2e-4) -> List[str]:
    """The embedder's parity rule: every component within rtol * max|want| of the reference (fp32 network,
    different but equivalent summation orders), and the direction -- what the cosine search consumes --
    within 1e-6 of cosine 1.  Returns a list of violations (empty = pass)."""
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    bad = []
    for i in range(want.shape[0]):
        scale = max(np.abs(want[i]).max(), 1e-30)
        err = np.abs(got[i] - want[i]).max() / scale
        cos = float(got[i] @ want[i] / max(np.linalg.norm(got[i]) * np.linalg.norm(want[i]), 1e-300))
        if not np.isfinite(err) or err > rtol or cos < 1 - 1e-6:
            bad.append(f"structure {i}: max rel err {err:.3e}, cosine {cos:.9f}")
    return bad
