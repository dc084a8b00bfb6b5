"""Host-side search engines over libfcsearch handles.

* ``shard_ranges``      -- contiguous row partition: shard g of G holds rows
  ``[g*ceil(N/G), min(N,(g+1)*ceil(N/G)))``; global id = local id + offset (the reference's
  ``I += i0``, dbsearch.py:238).
* ``plan_shards``       -- how many GPUs a single-process database uses (one for CATH scale, all for TED scale).
* ``LocalEngine``       -- ONE process, one host thread, one shard per chosen GPU (what the CLI user of
  ``merizo.py search -d cuda`` gets): a mirror of the library's shard group (csrc/fcs_group.cu) -- queries
  replicated, shards searched concurrently and asynchronously, per-shard key lists copied device-to-device
  (NVLink peer copies) to the first GPU and merged there by K5.
* ``DistributedEngine`` -- one process PER GPU under torchrun: each rank owns one shard; the
  per-rank ``[nq,k]`` packed key lists are exchanged with a single NCCL all-gather (8 bytes per
  entry over NVLink) and merged on every rank by the same kernel.  The key exchange is the only
  collective on the path.
* ``embed_partition`` / ``distributed_embed`` -- the step before the search under torchrun: the
  structures of a batch are independent, so each rank embeds a contiguous slice (balanced by
  sum of L^2, the embedder's cost) and ONE all-gather of the [n/world,128] fp32 embeddings
  (512 B per structure) gives every rank the replicated query matrix the search takes.

torch is used for device memory, streams and torch.distributed only.
"""
from __future__ import annotations

from typing import Callable, Iterable, List, Optional, Sequence, Tuple

import numpy as np

from . import native

DIM = native.DIM


def shard_ranges(n_rows: int, n_shards: int) -> List[Tuple[int, int]]:
    """Contiguous row ranges; trailing shards may be empty when n_rows < n_shards."""
    if n_rows < 0 or n_shards < 1:
        raise ValueError("n_rows >= 0 and n_shards >= 1 required")
    per = -(-n_rows // n_shards) if n_rows else 0
    return [(min(n_rows, g * per), min(n_rows, (g + 1) * per)) for g in range(n_shards)]


def encode_keys(scores: np.ndarray, ids: np.ndarray) -> np.ndarray:
    """Host mirror of the library's packed sort key (fcs_common.cuh make_key): order-preserving score word << 32 |
    (0xFFFFFFFF - id); key 0 = empty slot."""
    s = np.ascontiguousarray(scores, dtype=np.float32).view(np.uint32).astype(np.uint64)
    neg = (s & np.uint64(0x80000000)) != 0
    ordered = np.where(neg, (~s) & np.uint64(0xFFFFFFFF), s | np.uint64(0x80000000))
    ids = np.asarray(ids, dtype=np.int64)
    key = (ordered << np.uint64(32)) | (np.uint64(0xFFFFFFFF) - ids.astype(np.uint64))
    return np.where(ids < 0, np.uint64(0), key)


def decode_keys(keys: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    keys = np.asarray(keys, dtype=np.uint64)
    hi = (keys >> np.uint64(32)).astype(np.uint32)
    bits = np.where(hi & np.uint32(0x80000000), hi & np.uint32(0x7FFFFFFF), ~hi).astype(np.uint32)
    scores = bits.view(np.float32).copy()
    ids = (np.uint64(0xFFFFFFFF) - (keys & np.uint64(0xFFFFFFFF))).astype(np.int64)
    empty = keys == 0
    scores[empty] = -np.inf
    ids[empty] = -1
    return scores, ids


# ---- how many GPUs a single-process database uses -------------------------------------------------------------
MIN_ROWS_PER_SHARD = 4_000_000   # a shard below ~2 GB scans in < 0.3 ms: the key exchange + merge (~40 us) stops paying
HBM_FILL = 0.80                  # fraction of a GPU's free memory a shard may take


def plan_shards(n_rows: int, n_devices: int, bytes_per_row: int = 512, free_bytes: Optional[float] = None,
                min_rows_per_shard: int = MIN_ROWS_PER_SHARD) -> int:
    """Number of row shards (= GPUs) for a database of n_rows rows on a node with n_devices GPUs: as few as keep the
    per-shard scan well above the cross-GPU exchange cost, as many as memory needs.  A CATH-scale database (500 k rows,
    256 MB, 44 us per scan) stays on ONE GPU; the 365 M-row TED database takes all 8."""
    if n_devices < 1:
        raise ValueError("n_devices >= 1 required")
    by_size = max(1, n_rows // max(1, min_rows_per_shard))
    shards = min(n_devices, by_size)
    if free_bytes:
        need = -(-int(n_rows * bytes_per_row) // int(free_bytes * HBM_FILL))
        if need > n_devices:
            raise native.FcsError(native.ERR_NOMEM, f"{n_rows} rows x {bytes_per_row} B do not fit {n_devices} GPUs "
                                                    f"({free_bytes / 1e9:.0f} GB free each)")
        shards = max(shards, need)
    return max(1, min(shards, max(1, n_rows)))


class LocalEngine:
    """All visible GPUs (or ``devices``) from ONE process and ONE host thread: a thin mirror of the library's shard
    group (``fcs_group``, csrc/fcs_group.cu).  Queries are replicated, every shard searches on its own GPU and stream,
    the per-shard key lists are copied device-to-device (NVLink) to the first GPU and merged there; the host sees one
    synchronisation per search.  ``devices`` may repeat an ordinal (several shards on one GPU: the exchange + merge
    path on a single-GPU box).  Rows are fed block by block; every shard is fed by its own loader thread."""

    def __init__(self, n_rows: int, devices: Optional[Sequence[int]] = None, normalise_rows: bool = False,
                 keep_bf16: bool = False, has_lengths: bool = False, n_shards: Optional[int] = None):
        if devices is None:
            visible = native.device_count()
            if visible == 0:
                raise native.FcsError(native.ERR_CUDA, "no CUDA device visible")
            if n_shards is None:
                n_shards = plan_shards(int(n_rows), visible, 512 + (256 if keep_bf16 else 0) + (2 if has_lengths else 0),
                                       _free_bytes_per_gpu())
            devices = list(range(min(visible, n_shards)))
        devices = list(devices)
        if len(devices) == 0:
            raise native.FcsError(native.ERR_CUDA, "no CUDA device given")
        devices = devices[:max(1, min(len(devices), int(n_rows)))]
        self.n_rows = int(n_rows)
        self.devices = devices
        self.has_lengths = has_lengths
        self.group = native.Group(self.n_rows, devices, normalise_rows=normalise_rows, keep_bf16=keep_bf16, has_lengths=has_lengths)
        self.ranges = list(self.group.ranges)
        assert self.ranges == shard_ranges(self.n_rows, len(devices))

    @property
    def n_shards(self) -> int:
        return len(self.devices)

    def shard(self, index: int) -> native.Database:
        """Borrowed handle of one shard (timing / info)."""
        return self.group.shard(index)

    # -- loading -----------------------------------------------------------------------------
    def upload(self, row0: int, rows: np.ndarray, lengths: Optional[np.ndarray] = None) -> None:
        """Rows [row0, row0+len(rows)) of the GLOBAL matrix; the library splits them across the shards that own them."""
        self.group.upload(row0, rows, lengths)

    def upload_blocks(self, blocks: Iterable[np.ndarray], lengths: Optional[np.ndarray] = None,
                      progress: Optional[Callable[[int], None]] = None) -> None:
        """Consume a row-block iterator once (the reference's db_iterator, dbutil.py:33-35)."""
        i0 = 0
        for xb in blocks:
            self.upload(i0, np.asarray(xb), None if lengths is None else lengths[i0:i0 + xb.shape[0]])
            i0 += xb.shape[0]
            if progress:
                progress(i0)
        if i0 != self.n_rows:
            raise native.FcsError(native.ERR_INVALID, f"iterator yielded {i0} rows, database has {self.n_rows}")

    def upload_file(self, path: str, file_offset: int = 0) -> None:
        """The whole matrix from a file of headerless fp32 rows (the faiss flavour's *_raw_128d_norm.db, dbutil.py:28-30):
        every shard reads its own byte range with positional reads from its own thread."""
        self.group.upload_file(path, file_offset, 0, self.n_rows)

    def finalize(self) -> None:
        self.group.finalize()

    # -- search ------------------------------------------------------------------------------
    def search(self, q: np.ndarray, k: int, qlen: Optional[np.ndarray] = None, mincov: float = 0.0,
               qnorm: int = native.QNORM_NONE, mode: int = native.MODE_AUTO, kprime: int = 0, out=None):
        """(scores f32 [nq,k], ids i64 [nq,k]) as host arrays; exact; global ids.  `out=(scores, ids)` reuses the caller's arrays."""
        return self.group.search(q, k, qlen=qlen, mincov=mincov, qnorm=qnorm, mode=mode, kprime=kprime, out=out)

    def close(self) -> None:
        self.group.close()

    def __len__(self) -> int:
        return self.n_rows


def _free_bytes_per_gpu() -> Optional[float]:
    try:
        import torch

        if torch.cuda.is_available():
            return float(min(torch.cuda.mem_get_info(d)[0] for d in range(torch.cuda.device_count())))
    except Exception:
        pass
    return None


class DistributedEngine:
    """One rank per GPU (torchrun).  The world is laid out as R row shards x Q query groups (world = R*Q,
    rank = r*Q + j): rank (r, j) holds row shard r and searches query slice j, so a database that fits in fewer
    than `world` shards is replicated Q times and large batches are split instead of being repeated on every GPU.
    Q = 1 is pure row sharding (the only option at TED scale).  One all-gather of the packed keys, then the R
    lists of every query are merged on every rank.  ``local_search`` and ``merge`` are injectable so that the rank
    plumbing (partition, offsets, all-gather layout) is testable on CPU with the gloo backend."""

    def __init__(self, n_rows_global: int, rank: Optional[int] = None, world_size: Optional[int] = None,
                 device: Optional[int] = None, normalise_rows: bool = False, keep_bf16: bool = False,
                 has_lengths: bool = False, create_handle: bool = True, query_groups: int = 1):
        import torch.distributed as dist

        self.rank = dist.get_rank() if rank is None else rank
        self.world = dist.get_world_size() if world_size is None else world_size
        if query_groups < 1 or self.world % query_groups != 0:
            raise ValueError(f"query_groups={query_groups} must divide the world size {self.world}")
        self.q_groups = int(query_groups)
        self.row_shards = self.world // self.q_groups
        self.shard_index, self.q_index = divmod(self.rank, self.q_groups)
        self.n_rows_global = int(n_rows_global)
        self.ranges = shard_ranges(self.n_rows_global, self.row_shards)
        self.row0, self.row1 = self.ranges[self.shard_index]
        if device is None:
            # one rank per GPU: the launcher's LOCAL_RANK names this rank's device; never default every rank to GPU 0
            import os

            if "LOCAL_RANK" in os.environ:
                device = int(os.environ["LOCAL_RANK"])
            elif self.world == 1:
                device = 0
            elif create_handle:
                raise ValueError("DistributedEngine: pass device= (or launch with torchrun so that LOCAL_RANK is set)")
        self.device = device
        self._side = None
        self._pinned = None
        self._last_keys = None
        self._last_queue_max = None
        self.db: Optional[native.Database] = None
        if create_handle:
            self.db = native.Database(self.row1 - self.row0, device=self.device, id_offset=self.row0,
                                      normalise_rows=normalise_rows, keep_bf16=keep_bf16, has_lengths=has_lengths)

    @staticmethod
    def auto_query_groups(n_rows_global: int, world: int, bytes_per_row: int = 512 + 256, budget_bytes: float = 80e9) -> int:
        """Largest Q dividing `world` such that a row shard of N/(world/Q) rows fits `budget_bytes` per GPU."""
        best = 1
        for q in range(1, world + 1):
            if world % q == 0 and -(-n_rows_global // (world // q)) * bytes_per_row <= budget_bytes:
                best = q
        return best

    @property
    def dev_index(self) -> int:
        if self.device is None:
            raise ValueError("DistributedEngine: no device was given for this rank")
        return int(self.device)

    def query_slice(self, nq: int) -> Tuple[int, int, int]:
        """(first query, one-past-last, padded slice length) of this rank's query group."""
        per = -(-nq // self.q_groups)
        lo = min(nq, self.q_index * per)
        return lo, min(nq, lo + per), per

    def local_keys(self, q_dev, nq: int, k: int, qlen=None, rows: Optional[int] = None, **kw):
        """This rank's packed keys (torch int64 CUDA tensor viewing uint64 keys): `rows` rows (default nq), the first nq
        hold the shard's [nq,k] sorted lists.  With rows > nq, row `rows-1` carries the length of the search's exact-scan
        fallback queue in its first word (fcs_search_queue_len_to): it travels with the keys in the one all-gather."""
        import torch

        rows = nq if rows is None else rows
        dev = torch.device("cuda", self.dev_index)
        keys = torch.zeros((rows, k), dtype=torch.int64, device=dev)
        cur = torch.cuda.current_stream(dev)
        st = cur
        if cur.cuda_stream == 0:
            # the library reads stream 0 as "the handle's own stream": never launch on the legacy default stream,
            # use a side stream ordered after what produced the queries and before what consumes the keys
            if self._side is None:
                self._side = torch.cuda.Stream(dev)
            self._side.wait_stream(cur)
            st = self._side
        if nq > 0:
            self.db.search_device(q_dev.data_ptr(), nq, k, 0, 0, out_keys_ptr=keys.data_ptr(), stream=st.cuda_stream, qlen=qlen, **kw)
            if rows > nq:
                self.db.queue_len_to(keys[rows - 1].data_ptr(), stream=st.cuda_stream)
        if st is not cur:
            cur.wait_stream(st)
        return keys

    def search_host(self, q_host, k: int, **kw):
        """Host arrays in (the same queries on every rank), host arrays out: H2D copy from pinned memory, shard search,
        NCCL key all-gather, GPU merge, D2H copy into pinned memory -- ONE host synchronisation and ONE collective.  This is
        the end-to-end call of the one-rank-per-GPU deployment.  The returned arrays are views of pinned buffers owned by the
        engine (valid until the next call).  Exact on every rank: the length of every rank's exact-scan fallback queue rides
        along with the keys; if some rank had to defer part of its queue (more than ASYNC_FALLBACK_QUERIES queries failed
        their certificate: pathological data), every rank sees it after the synchronisation and the exchange + merge is
        repeated on the completed lists."""
        import torch

        dev = torch.device("cuda", self.dev_index)
        qh = torch.as_tensor(q_host, dtype=torch.float32).reshape(-1, DIM)
        nq = qh.shape[0]
        pin = self._pinned
        if pin is None or pin["q"].shape[0] < nq or pin["k"] != k:
            pin = self._pinned = {"q": torch.empty((nq, DIM), dtype=torch.float32).pin_memory(), "k": k,
                                  "sc": torch.empty((nq, k), dtype=torch.float32).pin_memory(),
                                  "ids": torch.empty((nq, k), dtype=torch.int64).pin_memory(),
                                  "flag": torch.zeros(1, dtype=torch.int64).pin_memory()}
        if qh.is_pinned():
            src = qh
        else:
            pin["q"][:nq].copy_(qh)
            src = pin["q"][:nq]
        q_dev = src.to(dev, non_blocking=True)
        cur = torch.cuda.current_stream(dev)
        for attempt in range(2):
            sc, ids = self.search(q_dev, k, **kw)
            pin["sc"][:nq].copy_(sc, non_blocking=True)
            pin["ids"][:nq].copy_(ids, non_blocking=True)
            if self._last_queue_max is not None:
                pin["flag"].copy_(self._last_queue_max, non_blocking=True)
            else:
                pin["flag"].zero_()
            cur.synchronize()
            if int(pin["flag"][0]) <= native.ASYNC_FALLBACK_QUERIES:
                break
            # some rank deferred part of its queue: every rank finishes its own (a no-op for most), then the (idempotent)
            # exchange + merge is repeated on the completed lists
            if self.db is not None:
                self.db.search_finish(cur.cuda_stream)
            kw = dict(kw, _reuse_local=True)
        return pin["sc"][:nq].numpy(), pin["ids"][:nq].numpy()

    def search(self, q_dev, k: int, local_search=None, merge=None, qlen=None, _reuse_local=False, **kw):
        """Replicated queries in, identical (scores, ids) on every rank out (torch tensors).  Asynchronous on the current
        stream: nothing here waits for the host.  ``self._last_queue_max`` (a 1-element device tensor, or None when the
        shard search is injected) is the longest exact-scan fallback queue over all ranks; callers that need the exactness
        guarantee under pathological data check it after their synchronisation (search_host does)."""
        import torch
        import torch.distributed as dist

        nq = q_dev.shape[0]
        lo, hi, per = self.query_slice(nq)
        q_mine = q_dev[lo:hi]
        ql_mine = None if qlen is None else np.asarray(qlen)[lo:hi]
        carry = local_search is None and self.db is not None  # one extra key row per rank carries the queue length
        rows = per + 1 if carry else per
        if _reuse_local and self._last_keys is not None:
            keys = self._last_keys  # the previous call's shard list, completed in place by search_finish
            if carry:
                keys[rows - 1].zero_()
        elif carry:
            keys = self.local_keys(q_mine, hi - lo, k, qlen=ql_mine, rows=rows, **kw)
        else:
            keys = local_search(q_mine, hi - lo, k) if hi - lo > 0 else torch.zeros((0, k), dtype=torch.int64, device=q_dev.device)
            if hi - lo < per:  # pad the slice: every rank contributes the same shape; key 0 = empty
                keys = torch.cat([keys, torch.zeros((per - (hi - lo), k), dtype=torch.int64, device=keys.device)], dim=0)
        self._last_keys = keys
        gathered = torch.empty((self.world, rows, k), dtype=keys.dtype, device=keys.device)
        # the one collective on the path: 8 B per entry (flat [world*rows, k] view: the layout gloo and nccl both accept)
        dist.all_gather_into_tensor(gathered.view(self.world * rows, k), keys.contiguous())
        # rank = r*Q + j  =>  gathered is [R][Q*rows][k]: R sorted lists for each of the Q*rows (padded) query rows
        lists = gathered.view(self.row_shards, self.q_groups * rows, k)
        if merge:
            sc, ids = merge(lists, k)
        else:
            sc = torch.empty((self.q_groups * rows, k), dtype=torch.float32, device=keys.device)
            ids = torch.empty((self.q_groups * rows, k), dtype=torch.int64, device=keys.device)
            st = torch.cuda.current_stream(keys.device)
            native.merge_topk(self.dev_index, lists.data_ptr(), self.row_shards, self.q_groups * rows, k, sc.data_ptr(),
                              ids.data_ptr(), stream=st.cuda_stream)
        self._last_queue_max = None
        if carry:
            # drop the carrier rows (merged garbage) and keep the longest queue length any rank reported
            self._last_queue_max = (gathered[:, rows - 1, 0] & 0xFFFFFFFF).max().reshape(1)
            if self.q_groups > 1:
                sc = sc.view(self.q_groups, rows, k)[:, :per].reshape(self.q_groups * per, k)
                ids = ids.view(self.q_groups, rows, k)[:, :per].reshape(self.q_groups * per, k)
        return sc[:nq], ids[:nq]


def embed_partition(lengths: Sequence[int], parts: int) -> List[Tuple[int, int]]:
    """Contiguous slices [lo, hi) of a batch of structures, one per rank, balanced by sum of L^2 (the edge kernel's
    cost is quadratic in the structure length; counts alone would leave the rank with the long chains behind)."""
    if parts < 1:
        raise ValueError("parts >= 1 required")
    n = len(lengths)
    cost = np.asarray(lengths, dtype=np.float64) ** 2
    cum = np.concatenate([[0.0], np.cumsum(cost)])
    total = cum[-1]
    bounds = [0]
    for p in range(1, parts):
        # first index whose prefix cost reaches p/parts of the total; never before the previous boundary
        b = int(np.searchsorted(cum, total * p / parts, side="left"))
        bounds.append(min(n, max(b, bounds[-1])))
    bounds.append(n)
    return [(bounds[p], bounds[p + 1]) for p in range(parts)]


def distributed_embed(structures: Sequence[np.ndarray], embed_fn: Callable, device=None):
    """Under torch.distributed (same `structures` on every rank): rank r embeds its slice with
    ``embed_fn(list_of_coords) -> [m,128]`` (torch tensor on `device`, or a numpy array), then one
    ``all_gather_into_tensor`` of the padded [per,128] blocks; returns the full [n,128] torch tensor, identical on
    every rank.  With one process it is just ``embed_fn(structures)``."""
    import torch
    import torch.distributed as dist

    n = len(structures)
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    lens = [int(np.asarray(c).reshape(-1, 3).shape[0]) for c in structures]
    parts = embed_partition(lens, world)
    lo, hi = parts[rank]
    mine = embed_fn(list(structures[lo:hi])) if hi > lo else np.zeros((0, DIM), dtype=np.float32)
    mine = torch.as_tensor(mine, dtype=torch.float32)
    if device is not None:
        mine = mine.to(device)
    if world == 1:
        return mine
    per = max(h - l for l, h in parts)
    block = torch.zeros((per, DIM), dtype=torch.float32, device=mine.device)
    block[: hi - lo] = mine
    gathered = torch.empty((world * per, DIM), dtype=torch.float32, device=mine.device)
    dist.all_gather_into_tensor(gathered, block)  # the only collective of the embedding step: 512 B per structure
    out = torch.empty((n, DIM), dtype=torch.float32, device=mine.device)
    for r, (l, h) in enumerate(parts):
        out[l:h] = gathered[r * per: r * per + (h - l)]
    return out
