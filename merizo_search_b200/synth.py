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
