"""ctypes binding of libfcsearch.so (include/fcsearch.h) -- thin, no compute, no fallback.

Every product entry point goes through here; if the CUDA library is missing or the call
fails, ``FcsError`` is raised -- there is no CPU path behind this module.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

from . import build as _build

DIM = 128
MAX_K = 2048

# error codes / flags / enums (mirrors include/fcsearch.h)
OK, ERR_INVALID, ERR_CUDA, ERR_STATE, ERR_UNSUPPORTED, ERR_NOMEM = 0, -1, -2, -3, -4, -5
DB_NORMALISE_ROWS, DB_KEEP_BF16, DB_HAS_LENGTHS = 1, 2, 4
QNORM_NONE, QNORM_COSINE, QNORM_L2 = 0, 1, 2
MODE_AUTO, MODE_GEMV, MODE_TC = 0, 1, 2
EMBED_MODE_FP32, EMBED_MODE_TC, EMBED_MODE_TC2, EMBED_MODE_TC3 = 0, 1, 2, 3

EXPORTS = [
    "fcs_version", "fcs_last_error", "fcs_device_count", "fcs_db_create", "fcs_db_upload",
    "fcs_db_upload_device", "fcs_db_finalize", "fcs_db_get_info", "fcs_db_destroy", "fcs_search",
    "fcs_search_device", "fcs_search_finish", "fcs_search_queue_len_to", "fcs_merge_topk", "fcs_get_timing", "fcs_set_profiling", "fcs_debug_tc_approx",
    "fcs_debug_tc_plan", "fcs_debug_tc_tile_of", "fcs_db_upload_file",
    "fcs_group_create", "fcs_group_upload", "fcs_group_upload_file", "fcs_group_finalize", "fcs_group_search",
    "fcs_group_get_info", "fcs_group_shard", "fcs_group_last_fallbacks", "fcs_group_destroy",
    # include/fcsembed.h
    "fcs_embedder_create", "fcs_embedder_destroy", "fcs_embed", "fcs_embed_to_device", "fcs_embed_get_timing",
    "fcs_embed_debug_layer", "fcs_embed_set_mode",
]


class FcsError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libfcsearch error {code}: {msg}")
        self.code = code


class Timing(C.Structure):
    _fields_ = [("last_search_ms", C.c_float), ("last_kernel_ms", C.c_float), ("last_mode", C.c_int32),
                ("last_launches", C.c_int32), ("last_tc_fallbacks", C.c_int32), ("last_rounds", C.c_int32)]


class Info(C.Structure):
    _fields_ = [("n_rows", C.c_int64), ("id_offset", C.c_int64), ("device", C.c_int32), ("flags", C.c_uint32),
                ("finalized", C.c_int32), ("sm_count", C.c_int32), ("bytes_fp32", C.c_uint64), ("bytes_bf16", C.c_uint64)]


class EgnnWeights(C.Structure):
    """fcs_egnn_weights: ten host pointers (fp32, nn.Linear layout), include/fcsembed.h."""
    FIELDS = ("edge_w1", "edge_b1", "edge_w2", "edge_b2", "gate_w", "gate_b", "node_w1", "node_b1", "node_w2", "node_b2")
    _fields_ = [(name, C.c_void_p) for name in FIELDS]


class EmbedTiming(C.Structure):
    _fields_ = [("last_ms", C.c_float), ("last_edge_ms", C.c_float), ("last_launches", C.c_int32),
                ("last_structures", C.c_int32), ("last_residues", C.c_int64), ("last_pairs", C.c_int64)]


_lib: Optional[C.CDLL] = None


def lib_path() -> str:
    return _build.LIB_PATH


def load() -> C.CDLL:
    """dlopen libfcsearch.so (building it first if the sources are newer) and type its exports."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB_PATH
    if not os.path.exists(path) or (not _build.is_fresh() and os.environ.get("FCS_NO_REBUILD") != "1"):
        try:
            _build.build()
        except Exception as exc:  # no silent fallback, and never a stale binary: the product needs THIS source built
            raise FcsError(ERR_STATE, f"libfcsearch.so is missing or stale and could not be (re)built: {exc}") from exc
    lib = C.CDLL(path)
    vp, i64, i32, u32, f32 = C.c_void_p, C.c_int64, C.c_int, C.c_uint32, C.c_float
    lib.fcs_version.restype = C.c_int
    lib.fcs_version.argtypes = []
    lib.fcs_last_error.restype = C.c_char_p
    lib.fcs_last_error.argtypes = []
    lib.fcs_device_count.argtypes = [C.POINTER(C.c_int)]
    lib.fcs_db_create.argtypes = [i32, i64, i32, i64, u32, C.POINTER(vp)]
    lib.fcs_db_upload.argtypes = [vp, i64, i64, vp, vp]
    lib.fcs_db_upload_device.argtypes = [vp, i64, i64, vp, vp]
    lib.fcs_db_upload_file.argtypes = [vp, i64, i64, C.c_char_p, i64]
    lib.fcs_group_create.argtypes = [C.POINTER(C.c_int), i32, i64, i32, u32, C.POINTER(vp)]
    lib.fcs_group_upload.argtypes = [vp, i64, i64, vp, vp]
    lib.fcs_group_upload_file.argtypes = [vp, C.c_char_p, i64, i64, i64]
    lib.fcs_group_finalize.argtypes = [vp]
    lib.fcs_group_search.argtypes = [vp, vp, i32, vp, f32, i32, i32, i32, i32, vp, vp]
    lib.fcs_group_get_info.argtypes = [vp, C.POINTER(C.c_int), vp, vp, i32]
    lib.fcs_group_shard.argtypes = [vp, i32, C.POINTER(vp)]
    lib.fcs_group_last_fallbacks.argtypes = [vp]
    lib.fcs_group_destroy.argtypes = [vp]
    lib.fcs_db_finalize.argtypes = [vp]
    lib.fcs_db_get_info.argtypes = [vp, C.POINTER(Info)]
    lib.fcs_db_destroy.argtypes = [vp]
    lib.fcs_search.argtypes = [vp, vp, i32, vp, f32, i32, i32, i32, i32, vp, vp]
    lib.fcs_search_device.argtypes = [vp, vp, i32, vp, f32, i32, i32, i32, i32, vp, vp, vp, vp]
    lib.fcs_search_finish.argtypes = [vp, vp, C.POINTER(C.c_int)]
    lib.fcs_search_queue_len_to.argtypes = [vp, vp, vp]
    lib.fcs_debug_tc_plan.argtypes = [i64, i32, i32, vp, i32]
    lib.fcs_debug_tc_tile_of.argtypes = [i64, i64, i64, i64]
    lib.fcs_merge_topk.argtypes = [i32, vp, i32, i32, i32, vp, vp, vp]
    lib.fcs_get_timing.argtypes = [vp, C.POINTER(Timing)]
    lib.fcs_debug_tc_approx.argtypes = [vp, vp, i32, i32, vp]
    lib.fcs_set_profiling.argtypes = [vp, i32]
    lib.fcs_embedder_create.argtypes = [i32, C.POINTER(EgnnWeights), i32, vp, i32, C.POINTER(vp)]
    lib.fcs_embedder_destroy.argtypes = [vp]
    lib.fcs_embed.argtypes = [vp, vp, vp, i32, vp]
    lib.fcs_embed_to_device.argtypes = [vp, vp, vp, i32, vp]
    lib.fcs_embed_get_timing.argtypes = [vp, C.POINTER(EmbedTiming)]
    lib.fcs_embed_debug_layer.argtypes = [vp, vp, i32, i32, vp, vp]
    lib.fcs_embed_set_mode.argtypes = [vp, i32]
    for name in EXPORTS:
        if name not in ("fcs_last_error",):
            getattr(lib, name).restype = C.c_int
    lib.fcs_last_error.restype = C.c_char_p
    lib.fcs_debug_tc_tile_of.restype = C.c_int64
    _lib = lib
    return lib


def _check(rc: int) -> None:
    if rc != OK:
        raise FcsError(rc, load().fcs_last_error().decode("utf-8", "replace"))


def device_count() -> int:
    n = C.c_int(0)
    _check(load().fcs_device_count(C.byref(n)))
    return n.value


_addressof, _char_from_buffer = C.addressof, C.c_char.from_buffer


def _np_ptr(a: Optional[np.ndarray]):
    """Address of an array's first byte, for a c_void_p parameter.  `ndarray.ctypes` costs 4 us per use (20 us per search
    call, a quarter of a single-query search); the buffer protocol gives the address in 0.8 us.  Read-only or empty arrays
    (np.memmap opened 'r', zero rows) take the slow way."""
    if a is None:
        return None
    try:
        return _addressof(_char_from_buffer(a))
    except (TypeError, ValueError, BufferError):
        return a.ctypes.data


def _out_arrays(out, nq: int, k: int):
    """Result arrays of a host-buffer search: fresh ones, or the caller's `out=(scores, ids)` after a shape/dtype check."""
    if out is None:
        return np.empty((nq, k), dtype=np.float32), np.empty((nq, k), dtype=np.int64)
    scores, ids = out
    if (scores.shape != (nq, k) or ids.shape != (nq, k) or scores.dtype != np.float32 or ids.dtype != np.int64
            or not scores.flags.c_contiguous or not ids.flags.c_contiguous):
        raise FcsError(ERR_INVALID, f"out must be C-contiguous (float32 [{nq},{k}], int64 [{nq},{k}])")
    return scores, ids


class Database:
    """One device-resident row shard (fcs_db).  Not re-entrant; do not share across forks."""

    def __init__(self, n_rows: int, device: int = 0, id_offset: int = 0, normalise_rows: bool = False,
                 keep_bf16: bool = False, has_lengths: bool = False):
        self._lib = load()
        self._h = C.c_void_p()
        flags = (DB_NORMALISE_ROWS if normalise_rows else 0) | (DB_KEEP_BF16 if keep_bf16 else 0) | \
                (DB_HAS_LENGTHS if has_lengths else 0)
        _check(self._lib.fcs_db_create(int(device), int(n_rows), DIM, int(id_offset), flags, C.byref(self._h)))
        self.n_rows, self.device, self.id_offset, self.flags = int(n_rows), int(device), int(id_offset), flags
        self.has_lengths = bool(has_lengths)

    # -- loading ---------------------------------------------------------------------------
    def upload(self, row0: int, rows: np.ndarray, lengths: Optional[np.ndarray] = None) -> None:
        rows = np.ascontiguousarray(rows, dtype=np.float32)
        if rows.ndim != 2 or rows.shape[1] != DIM:
            raise FcsError(ERR_INVALID, f"rows must be [n,{DIM}] float32, got {rows.shape}")
        lens = None if lengths is None else np.ascontiguousarray(lengths, dtype=np.int32)
        if lens is not None and lens.shape[0] != rows.shape[0]:
            raise FcsError(ERR_INVALID, "lengths and rows disagree on the row count")
        _check(self._lib.fcs_db_upload(self._h, int(row0), rows.shape[0], _np_ptr(rows), _np_ptr(lens)))

    def upload_file(self, row0: int, n: int, path: str, file_offset: int = 0) -> None:
        """Rows [row0, row0+n) from a file of headerless fp32 rows, starting at byte `file_offset`."""
        _check(self._lib.fcs_db_upload_file(self._h, int(row0), int(n), os.fsencode(path), int(file_offset)))

    def upload_device(self, row0: int, n: int, rows_ptr: int, lengths_ptr: Optional[int] = None) -> None:
        _check(self._lib.fcs_db_upload_device(self._h, int(row0), int(n), C.c_void_p(rows_ptr),
                                              C.c_void_p(lengths_ptr) if lengths_ptr else None))

    def finalize(self) -> None:
        _check(self._lib.fcs_db_finalize(self._h))

    # -- search ----------------------------------------------------------------------------
    def search(self, q: np.ndarray, k: int, qlen: Optional[np.ndarray] = None, mincov: float = 0.0,
               qnorm: int = QNORM_NONE, mode: int = MODE_AUTO, kprime: int = 0, out=None):
        """Host buffers in, host buffers out: (scores f32 [nq,k], ids i64 [nq,k]).  `out=(scores, ids)` reuses the caller's
        arrays (a loop that searches batch after batch avoids faulting in fresh result pages every call)."""
        q = np.ascontiguousarray(q, dtype=np.float32).reshape(-1, DIM)
        nq = q.shape[0]
        ql = None if qlen is None else np.ascontiguousarray(qlen, dtype=np.int32).reshape(-1)
        if ql is not None and ql.shape[0] != nq:
            raise FcsError(ERR_INVALID, "qlen must have one entry per query")
        scores, ids = _out_arrays(out, nq, int(k))
        _check(self._lib.fcs_search(self._h, _np_ptr(q), nq, _np_ptr(ql), float(mincov), int(k), int(qnorm), int(mode),
                                    int(kprime), _np_ptr(scores), _np_ptr(ids)))
        return scores, ids

    def search_device(self, q_ptr: int, nq: int, k: int, out_scores_ptr: int, out_ids_ptr: int,
                      out_keys_ptr: int = 0, qlen: Optional[np.ndarray] = None, mincov: float = 0.0,
                      qnorm: int = QNORM_NONE, mode: int = MODE_AUTO, kprime: int = 0, stream: int = 0) -> None:
        """Device pointers, asynchronous on `stream` (0 = the handle's own stream)."""
        ql = None if qlen is None else np.ascontiguousarray(qlen, dtype=np.int32).reshape(-1)
        _check(self._lib.fcs_search_device(
            self._h, C.c_void_p(q_ptr), int(nq), _np_ptr(ql), float(mincov), int(k), int(qnorm), int(mode), int(kprime),
            C.c_void_p(out_scores_ptr) if out_scores_ptr else None, C.c_void_p(out_ids_ptr) if out_ids_ptr else None,
            C.c_void_p(out_keys_ptr) if out_keys_ptr else None, C.c_void_p(stream) if stream else None))

    def search_finish(self, stream: int = 0) -> int:
        """Synchronise `stream` and complete whatever the last asynchronous tensor-core search deferred (a fallback
        queue longer than FCS_ASYNC_FALLBACK_QUERIES); returns the length of that queue."""
        n = C.c_int(0)
        _check(self._lib.fcs_search_finish(self._h, C.c_void_p(stream) if stream else None, C.byref(n)))
        return n.value

    def queue_len_to(self, dst_ptr: int, stream: int = 0) -> None:
        """Enqueue a copy of the last asynchronous search's fallback-queue length (uint32) to device memory at dst_ptr."""
        _check(self._lib.fcs_search_queue_len_to(self._h, C.c_void_p(dst_ptr), C.c_void_p(stream) if stream else None))

    def debug_tc_approx(self, q: np.ndarray, qnorm: int = QNORM_NONE) -> np.ndarray:
        """Test hook: bf16 tensor-core scores [nq, n_rows] (shards of <= 4096 rows)."""
        q = np.ascontiguousarray(q, dtype=np.float32).reshape(-1, DIM)
        out = np.empty((q.shape[0], self.n_rows), dtype=np.float32)
        _check(self._lib.fcs_debug_tc_approx(self._h, _np_ptr(q), q.shape[0], int(qnorm), _np_ptr(out)))
        return out

    def set_profiling(self, enable: bool) -> None:
        _check(self._lib.fcs_set_profiling(self._h, 1 if enable else 0))

    def timing(self) -> Timing:
        t = Timing()
        _check(self._lib.fcs_get_timing(self._h, C.byref(t)))
        return t

    def info(self) -> Info:
        i = Info()
        _check(self._lib.fcs_db_get_info(self._h, C.byref(i)))
        return i

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.fcs_db_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _BorrowedShard(Database):
    """A shard handle owned by a Group (timing / info only; never destroyed from here)."""

    def __init__(self, lib, handle, n_rows, device, id_offset, flags):
        self._lib, self._h = lib, handle
        self.n_rows, self.device, self.id_offset, self.flags = n_rows, device, id_offset, flags
        self.has_lengths = bool(flags & DB_HAS_LENGTHS)

    def close(self) -> None:
        self._h = C.c_void_p()


class Group:
    """Several row shards behind one handle (fcs_group): one host thread drives all of them; key lists are exchanged
    device-to-device and merged on the first GPU.  `devices[s]` = GPU of shard s (ordinals may repeat)."""

    def __init__(self, n_rows: int, devices, normalise_rows: bool = False, keep_bf16: bool = False, has_lengths: bool = False):
        self._lib = load()
        self._h = C.c_void_p()
        devices = [int(d) for d in devices]
        flags = (DB_NORMALISE_ROWS if normalise_rows else 0) | (DB_KEEP_BF16 if keep_bf16 else 0) | \
                (DB_HAS_LENGTHS if has_lengths else 0)
        arr = (C.c_int * len(devices))(*devices)
        _check(self._lib.fcs_group_create(arr, len(devices), int(n_rows), DIM, flags, C.byref(self._h)))
        self.n_rows, self.devices, self.flags = int(n_rows), devices, flags
        bounds = np.zeros(len(devices) + 1, dtype=np.int64)
        n = C.c_int(0)
        _check(self._lib.fcs_group_get_info(self._h, C.byref(n), _np_ptr(bounds), None, len(devices)))
        self.bounds = [int(b) for b in bounds]
        self.ranges = [(self.bounds[i], self.bounds[i + 1]) for i in range(len(devices))]

    def shard(self, index: int) -> Database:
        h = C.c_void_p()
        _check(self._lib.fcs_group_shard(self._h, int(index), C.byref(h)))
        r0, r1 = self.ranges[index]
        return _BorrowedShard(self._lib, h, r1 - r0, self.devices[index], r0, self.flags)

    def upload(self, row0: int, rows: np.ndarray, lengths: Optional[np.ndarray] = None) -> None:
        rows = np.ascontiguousarray(rows, dtype=np.float32)
        if rows.ndim != 2 or rows.shape[1] != DIM:
            raise FcsError(ERR_INVALID, f"rows must be [n,{DIM}] float32, got {rows.shape}")
        lens = None if lengths is None else np.ascontiguousarray(lengths, dtype=np.int32)
        if lens is not None and lens.shape[0] != rows.shape[0]:
            raise FcsError(ERR_INVALID, "lengths and rows disagree on the row count")
        _check(self._lib.fcs_group_upload(self._h, int(row0), rows.shape[0], _np_ptr(rows), _np_ptr(lens)))

    def upload_file(self, path: str, file_offset: int, row0: int, n: int) -> None:
        _check(self._lib.fcs_group_upload_file(self._h, os.fsencode(path), int(file_offset), int(row0), int(n)))

    def finalize(self) -> None:
        _check(self._lib.fcs_group_finalize(self._h))

    def search(self, q: np.ndarray, k: int, qlen: Optional[np.ndarray] = None, mincov: float = 0.0,
               qnorm: int = QNORM_NONE, mode: int = MODE_AUTO, kprime: int = 0, out=None):
        q = np.ascontiguousarray(q, dtype=np.float32).reshape(-1, DIM)
        nq = q.shape[0]
        ql = None if qlen is None else np.ascontiguousarray(qlen, dtype=np.int32).reshape(-1)
        if ql is not None and ql.shape[0] != nq:
            raise FcsError(ERR_INVALID, "qlen must have one entry per query")
        scores, ids = _out_arrays(out, nq, int(k))
        _check(self._lib.fcs_group_search(self._h, _np_ptr(q), nq, _np_ptr(ql), float(mincov), int(k), int(qnorm), int(mode),
                                          int(kprime), _np_ptr(scores), _np_ptr(ids)))
        return scores, ids

    def last_fallbacks(self) -> int:
        return int(self._lib.fcs_group_last_fallbacks(self._h))

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.fcs_group_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


ASYNC_FALLBACK_QUERIES = 32

TC_PLAN_FIELDS = ("tiles", "j0", "stride", "comp_t", "first", "rank", "partition")


def debug_tc_plan(n_rows: int, kprime: int = 0, nq: int = 512, max_rounds: int = 16):
    """Test hook: the rounds the tensor-core path runs for a shard of n_rows rows and nq queries (list of dicts, TC_PLAN_FIELDS)."""
    out = np.zeros((max_rounds, len(TC_PLAN_FIELDS)), dtype=np.int64)
    n = load().fcs_debug_tc_plan(int(n_rows), int(kprime), int(nq), _np_ptr(out), int(max_rounds))
    if n < 0:
        _check(n)
    return [dict(zip(TC_PLAN_FIELDS, (int(v) for v in out[i]))) for i in range(n)]


def debug_tc_tile_of(j0: int, stride: int, comp_t: int, idx: int) -> int:
    return int(load().fcs_debug_tc_tile_of(int(j0), int(stride), int(comp_t), int(idx)))


def merge_topk(device: int, keys_ptr: int, n_lists: int, nq: int, k: int, out_scores_ptr: int, out_ids_ptr: int,
               stream: int = 0) -> None:
    _check(load().fcs_merge_topk(int(device), C.c_void_p(keys_ptr), int(n_lists), int(nq), int(k),
                                 C.c_void_p(out_scores_ptr), C.c_void_p(out_ids_ptr), C.c_void_p(stream) if stream else None))


# state_dict key suffix of every fcs_egnn_weights field (encode_ca_egnn.<layer>.<suffix>) and its shape
EGNN_KEYS = {
    "edge_w1": ("edge_mlp.0.weight", (514, 257)), "edge_b1": ("edge_mlp.0.bias", (514,)),
    "edge_w2": ("edge_mlp.2.weight", (256, 514)), "edge_b2": ("edge_mlp.2.bias", (256,)),
    "gate_w": ("edge_gate.0.weight", (1, 256)), "gate_b": ("edge_gate.0.bias", (1,)),
    "node_w1": ("node_mlp.0.weight", (256, 384)), "node_b1": ("node_mlp.0.bias", (256,)),
    "node_w2": ("node_mlp.2.weight", (128, 256)), "node_b2": ("node_mlp.2.bias", (128,)),
}


class Embedder:
    """Device-resident FoldClassNet(128) weights + the batched forward (fcs_embedder).  Not re-entrant."""

    def __init__(self, layers, pe: np.ndarray, device: int = 0):
        """layers: list (one per EGNN layer) of dicts field name -> fp32 array (EGNN_KEYS shapes);
        pe: the positional table [max_len, 128]."""
        self._lib = load()
        self._h = C.c_void_p()
        pe = np.ascontiguousarray(pe, dtype=np.float32).reshape(-1, DIM)
        keep = []  # the arrays must outlive the create call only (the library copies them to the device)
        arr = (EgnnWeights * len(layers))()
        for i, layer in enumerate(layers):
            for name in EgnnWeights.FIELDS:
                a = np.ascontiguousarray(layer[name], dtype=np.float32)
                if a.shape != EGNN_KEYS[name][1]:
                    raise FcsError(ERR_INVALID, f"layer {i}: {name} has shape {a.shape}, expected {EGNN_KEYS[name][1]}")
                keep.append(a)
                setattr(arr[i], name, a.ctypes.data_as(C.c_void_p))
        _check(self._lib.fcs_embedder_create(int(device), arr, len(layers), _np_ptr(pe), int(pe.shape[0]), C.byref(self._h)))
        self.device, self.n_layers, self.max_len = int(device), len(layers), int(pe.shape[0])

    @staticmethod
    def _pack(structures):
        lens = [int(np.asarray(c).reshape(-1, 3).shape[0]) for c in structures]
        offsets = np.zeros(len(lens) + 1, dtype=np.int64)
        offsets[1:] = np.cumsum(lens)
        if len(lens):
            coords = np.concatenate([np.asarray(c, dtype=np.float32).reshape(-1, 3) for c in structures])
        else:
            coords = np.zeros((0, 3), dtype=np.float32)
        return np.ascontiguousarray(coords, dtype=np.float32), offsets

    def embed(self, structures) -> np.ndarray:
        """list of [L,3] C-alpha traces -> host array [n,128]."""
        coords, offsets = self._pack(structures)
        return self.embed_packed(coords, offsets)

    def embed_packed(self, coords: np.ndarray, offsets: np.ndarray) -> np.ndarray:
        coords = np.ascontiguousarray(coords, dtype=np.float32).reshape(-1, 3)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        n = offsets.shape[0] - 1
        out = np.empty((n, DIM), dtype=np.float32)
        _check(self._lib.fcs_embed(self._h, _np_ptr(coords), _np_ptr(offsets), n, _np_ptr(out)))
        return out

    def embed_packed_to_device(self, coords: np.ndarray, offsets: np.ndarray, out_ptr: int) -> None:
        """Same, the [n,128] fp32 result is written to device memory at `out_ptr` (complete on return)."""
        coords = np.ascontiguousarray(coords, dtype=np.float32).reshape(-1, 3)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        _check(self._lib.fcs_embed_to_device(self._h, _np_ptr(coords), _np_ptr(offsets), offsets.shape[0] - 1, C.c_void_p(out_ptr)))

    def debug_layer(self, coords: np.ndarray, layer: int):
        """Test hook: (node features after `layer` [L,128], summed messages of that layer [L,256])."""
        coords = np.ascontiguousarray(coords, dtype=np.float32).reshape(-1, 3)
        L = coords.shape[0]
        feats = np.empty((L, DIM), dtype=np.float32)
        msgs = np.empty((L, 256), dtype=np.float32)
        _check(self._lib.fcs_embed_debug_layer(self._h, _np_ptr(coords), L, int(layer), _np_ptr(feats), _np_ptr(msgs)))
        return feats, msgs

    def set_mode(self, mode: int) -> None:
        """EMBED_MODE_FP32 (fp32 FMA pipe), EMBED_MODE_TC (round-1 tcgen05 kernel), EMBED_MODE_TC2 / EMBED_MODE_TC3 (tcgen05 with
        dedicated epilogue warps and two accumulator buffers, 8 / 16 generator warps; TC3 is the library's default)."""
        _check(self._lib.fcs_embed_set_mode(self._h, int(mode)))

    def timing(self) -> EmbedTiming:
        t = EmbedTiming()
        _check(self._lib.fcs_embed_get_timing(self._h, C.byref(t)))
        return t

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.fcs_embedder_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
