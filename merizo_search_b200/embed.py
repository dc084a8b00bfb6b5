"""Drop-in for the reference's query embedder.

The reference builds ``FoldClassNet(128)`` in ``network_setup`` (dbsearch.py:35-45) and calls it ONE
structure at a time (``network(query_input)``, dbsearch.py:97-98 and 287-301); every call materialises
[L,L,514] tensors and the embeddings go through the host before the search (dbsearch.py:316).

``FoldClassEmbedder`` holds the same weights on the GPU behind libfcsearch (``fcs_embedder``,
include/fcsembed.h; hand-written sm_100a kernels, merizo_search_b200/csrc/fcs_embed.cu):

  * ``embedder(x)`` with ``x`` a torch tensor [1,L,3] returns a [1,128] tensor -- the call contract of the
    reference module, so ``dbsearch()`` / ``dbsearch_faiss()`` run unchanged with it as ``network``;
  * ``embed_structures(list_of_coords)`` embeds a whole ragged batch in one call (what
    ``faiss_driver.embed_queries`` uses), ``embed_structures_device`` leaves the [n,128] matrix in HBM.

No CPU path: without the library or a GPU the constructor raises ``native.FcsError``.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np

from . import native

WIDTH = native.DIM
N_LAYERS = 2  # FoldClassNet: two EGNN layers (nndef_fold_egnn_embed.py:45-47)


def layers_from_state_dict(state_dict: Dict[str, object], n_layers: int = N_LAYERS) -> List[Dict[str, np.ndarray]]:
    """``encode_ca_egnn.<l>.<edge_mlp.0.weight ...>`` -> the per-layer dicts native.Embedder takes."""
    def arr(key):
        if key not in state_dict:
            raise KeyError(f"state_dict has no '{key}' (expected a FoldClassNet(128) checkpoint)")
        v = state_dict[key]
        if hasattr(v, "detach"):
            v = v.detach().to("cpu").float().numpy()
        return np.ascontiguousarray(v, dtype=np.float32)

    return [{field: arr(f"encode_ca_egnn.{l}.{suffix}") for field, (suffix, _shape) in native.EGNN_KEYS.items()}
            for l in range(n_layers)]


def positional_table_from_state_dict(state_dict: Dict[str, object]) -> np.ndarray:
    """``posenc_as.pe`` ([1, max_len, 128]; a registered buffer, so it is part of the checkpoint)."""
    v = state_dict["posenc_as.pe"]
    if hasattr(v, "detach"):
        v = v.detach().to("cpu").float().numpy()
    return np.ascontiguousarray(v, dtype=np.float32).reshape(-1, WIDTH)


class FoldClassEmbedder:
    """Callable like the reference's ``FoldClassNet`` instance; batched and device-resident underneath."""

    def __init__(self, state_dict: Dict[str, object], device=0):
        idx = _device_index(device)
        self._emb = native.Embedder(layers_from_state_dict(state_dict), positional_table_from_state_dict(state_dict), device=idx)
        self.device_index = idx
        self.width = WIDTH
        self.max_len = self._emb.max_len

    # ---- construction helpers ---------------------------------------------------------------
    @classmethod
    def from_network(cls, network, device=0) -> "FoldClassEmbedder":
        """From the reference's torch module (after its load_state_dict, dbsearch.py:43)."""
        return cls(network.state_dict(), device)

    @classmethod
    def from_file(cls, path: str, device=0) -> "FoldClassEmbedder":
        """From FINAL_foldclass_model.pt (a plain state_dict, dbsearch.py:43)."""
        import torch

        return cls(torch.load(path, map_location="cpu"), device)

    # ---- torch.nn.Module look-alikes the reference's drivers touch ------------------------------
    def eval(self):
        return self

    def to(self, *_args, **_kwargs):
        return self

    def __call__(self, x):
        """x: torch tensor [B,L,3] (the reference always passes B=1) -> torch tensor [B,128] on x's device."""
        import torch

        if not hasattr(x, "detach"):
            x = torch.as_tensor(np.asarray(x, dtype=np.float32))
        if x.dim() != 3 or x.shape[-1] != 3:
            raise ValueError(f"expected coordinates of shape [B,L,3], got {tuple(x.shape)}")
        coords = x.detach().to("cpu", torch.float32).contiguous().numpy()
        out = self.embed_structures([coords[b] for b in range(coords.shape[0])])
        return torch.from_numpy(out).to(x.device)

    forward = __call__

    # ---- batched API ------------------------------------------------------------------------------
    def embed_structures(self, structures: Sequence[np.ndarray]) -> np.ndarray:
        """list of [L,3] C-alpha traces (ragged) -> host array [n,128] fp32."""
        return self._emb.embed(structures)

    def embed_structures_device(self, structures: Sequence[np.ndarray]):
        """Same, the result stays in HBM as a torch CUDA tensor [n,128] (feed it to Database.search_device)."""
        import torch

        coords, offsets = native.Embedder._pack(structures)
        out = torch.empty((offsets.shape[0] - 1, WIDTH), dtype=torch.float32, device=torch.device("cuda", self.device_index))
        torch.cuda.current_stream(out.device).synchronize()  # the library writes on its own stream
        self._emb.embed_packed_to_device(coords, offsets, out.data_ptr())
        return out

    def timing(self) -> native.EmbedTiming:
        return self._emb.timing()

    def close(self) -> None:
        self._emb.close()


def _device_index(device) -> int:
    if isinstance(device, int):
        return device
    s = str(device)
    if s.startswith("cuda"):
        return int(s.split(":")[1]) if ":" in s else 0
    raise native.FcsError(native.ERR_INVALID, f"FoldClassEmbedder needs a CUDA device, got '{s}' (there is no CPU path)")


def wrap_network_setup(reference_network_setup):
    """Decorator for the reference's ``network_setup(threads, device)`` (dbsearch.py:35-45): the torch module it
    builds (and loads the checkpoint into) is replaced by a FoldClassEmbedder holding the same weights whenever
    the device is a CUDA device; otherwise the reference module is returned untouched."""
    def network_setup(*args, **kwargs):
        network, device = reference_network_setup(*args, **kwargs)
        if str(device).startswith("cuda"):
            network = FoldClassEmbedder.from_network(network, device)
        return network, device

    network_setup.__wrapped__ = reference_network_setup
    return network_setup
