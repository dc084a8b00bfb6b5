"""Test helper: merge packed key lists on the host (numpy).  Stands in for the CUDA merge kernel in the CPU-only
(gloo) tests of the multi-rank plumbing; deliberately NOT part of the product package."""
import numpy as np

from merizo_search_b200.engine import decode_keys


def merge_keys_host(keys: np.ndarray, k: int):
    """keys [n_lists, nq, k] (uint64, bigger = better, 0 = empty) -> (scores [nq,k], ids [nq,k])."""
    keys = np.asarray(keys, dtype=np.uint64)
    n_lists, nq, kk = keys.shape
    flat = np.transpose(keys, (1, 0, 2)).reshape(nq, n_lists * kk)
    top = np.sort(flat, axis=1)[:, ::-1][:, :k]  # descending, unsigned
    return decode_keys(top)
