"""One-time converted layout of a ``.pt`` database's ``.index`` file.

The reference pickles ``list[(path: str, ca: np.float32[L,3], seq: str)]`` with one numpy array per domain
(makedb.py:68-91) and ``read_database`` unpickles all N of them on every start and then walks the list in Python
to get the lengths (dbsearch.py:53-58): for a CATH-scale database that parse is seconds of pure host time before
the first query, although a search only ever touches the k records of its hits (dbsearch.py:124, 157;
dbsearch_fulllength.py:381-462).

``convert_index(db_name)`` rewrites that pickle ONCE into flat arrays next to it (``<db>.index.fcs/``):

    names.npy  uint8 [sum len(name)]   names_off.npy int64 [N+1]
    seq.npy    uint8 [sum len(seq)]    seq_off.npy   int64 [N+1]
    coords.npy float32 [sum L, 3]      coords_off.npy int64 [N+1]
    meta.json  {"n": N, "source_size": bytes, "source_mtime_ns": ..., "version": 1}

``LazyIndex`` maps them (``np.load(mmap_mode="r")``: opening is O(1)) and answers what the reference asks of the
list -- ``len()``, ``index[i]`` with anything that has ``__index__`` (torch 0-d tensors included), iteration -- by
materialising ``(name, coords, seq)`` per access; ``lengths()`` is one vectorised difference of offsets.
``read_database`` picks a sidecar up automatically when it is present and fresh (same size and mtime of the pickle
it was made from) and otherwise keeps the reference behaviour; ``FCS_INDEX_CACHE=1`` makes it write the sidecar after
the first slow load.  Host-side only: no GPU, no change to the on-disk formats of the reference.
"""
from __future__ import annotations

import json
import operator
import os
import pickle
from typing import List, Optional, Sequence, Tuple

import numpy as np

VERSION = 1
SUFFIX = ".index.fcs"
_ARRAYS = ("names", "names_off", "seq", "seq_off", "coords", "coords_off")


def sidecar_dir(db_name: str) -> str:
    return db_name + SUFFIX


def _source_stamp(index_path: str) -> dict:
    st = os.stat(index_path)
    return {"source_size": int(st.st_size), "source_mtime_ns": int(st.st_mtime_ns)}


def write_sidecar(out_dir: str, entries: Sequence[Tuple[str, np.ndarray, str]], stamp: Optional[dict] = None) -> str:
    """Flatten ``entries`` (the unpickled list) into ``out_dir``.  Written to a temporary directory and renamed, so a
    reader never sees a half-written sidecar."""
    n = len(entries)
    names_off = np.zeros(n + 1, dtype=np.int64)
    seq_off = np.zeros(n + 1, dtype=np.int64)
    coords_off = np.zeros(n + 1, dtype=np.int64)
    name_bytes: List[bytes] = []
    seq_bytes: List[bytes] = []
    coords: List[np.ndarray] = []
    for i, (name, ca, seq) in enumerate(entries):
        nb = str(name).encode("utf-8")
        sb = str(seq).encode("ascii")
        c = np.ascontiguousarray(ca, dtype=np.float32).reshape(-1, 3)
        name_bytes.append(nb)
        seq_bytes.append(sb)
        coords.append(c)
        names_off[i + 1] = names_off[i] + len(nb)
        seq_off[i + 1] = seq_off[i] + len(sb)
        coords_off[i + 1] = coords_off[i] + c.shape[0]
    arrays = {
        "names": np.frombuffer(b"".join(name_bytes), dtype=np.uint8),
        "names_off": names_off,
        "seq": np.frombuffer(b"".join(seq_bytes), dtype=np.uint8),
        "seq_off": seq_off,
        "coords": np.concatenate(coords) if coords else np.zeros((0, 3), np.float32),
        "coords_off": coords_off,
    }
    tmp = out_dir + f".tmp{os.getpid()}"
    os.makedirs(tmp, exist_ok=True)
    for key, arr in arrays.items():
        np.save(os.path.join(tmp, key + ".npy"), arr)
    meta = {"n": n, "version": VERSION}
    meta.update(stamp or {})
    with open(os.path.join(tmp, "meta.json"), "w") as fh:
        json.dump(meta, fh)
    if os.path.isdir(out_dir):  # replace an older sidecar
        for f in os.listdir(out_dir):
            os.unlink(os.path.join(out_dir, f))
        os.rmdir(out_dir)
    os.rename(tmp, out_dir)
    return out_dir


def convert_index(db_name: str, entries: Optional[Sequence] = None) -> str:
    """``<db_name>.index`` (the reference's pickle) -> ``<db_name>.index.fcs/``.  Returns the sidecar directory."""
    index_path = db_name + ".index"
    if entries is None:
        with open(index_path, "rb") as fh:
            entries = pickle.load(fh)
    return write_sidecar(sidecar_dir(db_name), entries, _source_stamp(index_path))


class LazyIndex:
    """Stands in for the unpickled ``list[(name, coords, seq)]`` of a ``.pt`` database."""

    def __init__(self, directory: str):
        with open(os.path.join(directory, "meta.json")) as fh:
            self.meta = json.load(fh)
        if self.meta.get("version") != VERSION:
            raise ValueError(f"{directory}: sidecar version {self.meta.get('version')} != {VERSION}")
        self.directory = directory
        a = {key: np.load(os.path.join(directory, key + ".npy"), mmap_mode="r") for key in _ARRAYS}
        self._names, self._names_off = a["names"], a["names_off"]
        self._seq, self._seq_off = a["seq"], a["seq_off"]
        self._coords, self._coords_off = a["coords"], a["coords_off"]
        self._n = int(self.meta["n"])
        if not (self._names_off.shape[0] == self._seq_off.shape[0] == self._coords_off.shape[0] == self._n + 1):
            raise ValueError(f"{directory}: offset tables disagree with n={self._n}")

    def __len__(self) -> int:
        return self._n

    def _item(self, i: int):
        name = bytes(self._names[self._names_off[i]:self._names_off[i + 1]]).decode("utf-8")
        seq = bytes(self._seq[self._seq_off[i]:self._seq_off[i + 1]]).decode("ascii")
        coords = np.array(self._coords[self._coords_off[i]:self._coords_off[i + 1]], dtype=np.float32)  # a private copy
        return name, coords, seq

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self._item(j) for j in range(*i.indices(self._n))]
        i = operator.index(i)  # ints, numpy integers, torch 0-d tensors (dbsearch.py:124 indexes with a tensor element)
        if i < 0:
            i += self._n
        if not 0 <= i < self._n:
            raise IndexError("list index out of range")
        return self._item(i)

    def __iter__(self):
        for i in range(self._n):
            yield self._item(i)

    def lengths(self) -> np.ndarray:
        """len(seq) of every domain (dbsearch.py:58), without touching the records."""
        return np.diff(np.asarray(self._seq_off)).astype(np.int32)

    def names(self, ids: Sequence[int]) -> List[str]:
        return [bytes(self._names[self._names_off[i]:self._names_off[i + 1]]).decode("utf-8") for i in map(operator.index, ids)]


def open_if_fresh(db_name: str) -> Optional[LazyIndex]:
    """The sidecar of ``<db_name>.index`` if there is one and it was made from exactly this pickle, else None."""
    d = sidecar_dir(db_name)
    index_path = db_name + ".index"
    if not (os.path.isdir(d) and os.path.exists(os.path.join(d, "meta.json")) and os.path.exists(index_path)):
        return None
    try:
        lazy = LazyIndex(d)
    except (OSError, ValueError, KeyError, json.JSONDecodeError):
        return None
    stamp = _source_stamp(index_path)
    if any(lazy.meta.get(k) != v for k, v in stamp.items()):
        return None  # the pickle changed after the conversion
    return lazy


def load_index(db_name: str):
    """(index object, lengths int32 [N]).  Fresh sidecar -> LazyIndex (O(1) open); otherwise the reference's own
    ``pickle.load`` + length walk (dbsearch.py:53-58), writing a sidecar afterwards when ``FCS_INDEX_CACHE=1``."""
    lazy = open_if_fresh(db_name)
    if lazy is not None:
        return lazy, lazy.lengths()
    with open(db_name + ".index", "rb") as fh:
        entries = pickle.load(fh)
    lengths = np.asarray([len(t[2]) for t in entries], dtype=np.int32)
    if os.environ.get("FCS_INDEX_CACHE") == "1":
        try:
            convert_index(db_name, entries)
        except OSError:
            pass  # read-only database directory: keep the reference behaviour
    return entries, lengths


if __name__ == "__main__":  # python -m merizo_search_b200.dbindex <db basename> [...]
    import sys
    import time

    for base in sys.argv[1:]:
        t0 = time.time()
        out = convert_index(base)
        print(f"{base}.index -> {out} ({len(LazyIndex(out))} domains, {time.time() - t0:.1f} s)")
