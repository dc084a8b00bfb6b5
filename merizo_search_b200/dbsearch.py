"""Drop-in replacements for the search hot path of merizo_search/programs/Foldclass/dbsearch.py.

Same names, argument meaning, return types and error behaviour as the reference callables, so the
reference's own drivers (``dbsearch``, ``run_dbsearch``, ``multi_domain_search``) keep working:

  read_database(db_name, device)                          reference dbsearch.py:48-72
  search_query_against_db(query_dict, target_dict, ...)   reference dbsearch.py:75-81
  knn_exact(xq, db_iterator, k, ...)                      reference dbsearch.py:213-248 (knn_exact_faiss)
  dbsearch_faiss(queries, target_dict, ...)               reference dbsearch.py:203-472 (see faiss_driver.py)
  install(reference_module)                               splice the above into the reference module (and wrap its
                                                          network_setup so CUDA runs get the batched embedder, embed.py)

The arithmetic runs in libfcsearch.so (hand-written sm_100a CUDA) on device-resident, row-sharded
databases; there is no CPU path here -- without the library or a GPU these functions raise.
"""
from __future__ import annotations

import logging
import os
import sys
import time
from typing import Dict, Iterable, Optional, Tuple

import numpy as np

from . import dbindex, native, serve
from .engine import LocalEngine

logger = logging.getLogger(__name__)

DIM = native.DIM
_RESIDENT: Dict[Tuple[str, str], "ResidentDatabase"] = {}  # (abs path, flavour) -> loaded database
MAX_RESIDENT = int(os.environ.get("FCS_MAX_RESIDENT", "2"))  # databases kept in HBM per process (least recently used goes)


def _file_stamp(path: str) -> Tuple[int, int]:
    st = os.stat(path)
    return (st.st_size, st.st_mtime_ns)


def _resident_get(key: Tuple[str, str], path: str) -> Optional["ResidentDatabase"]:
    """Cached database for `key` if the file behind it is still the one that was loaded (size + mtime): a database
    rebuilt in the same process (createdb, then search) is reloaded instead of being served stale from HBM."""
    r = _RESIDENT.get(key)
    if r is None:
        return None
    if r.stamp is not None and r.stamp != _file_stamp(path):
        r.engine.close()
        del _RESIDENT[key]
        return None
    _RESIDENT[key] = _RESIDENT.pop(key)  # most recently used last
    return r


def _resident_put(key: Tuple[str, str], resident: "ResidentDatabase", path: str) -> None:
    resident.stamp = _file_stamp(path)
    _RESIDENT[key] = resident
    while len(_RESIDENT) > max(1, MAX_RESIDENT):  # evict the least recently used: two TED-scale databases do not fit
        old_key = next(iter(_RESIDENT))
        _RESIDENT.pop(old_key).engine.close()


def _want_bf16() -> bool:
    """The bf16 operand images for the tcgen05 path cost +50 % HBM; on by default, FCS_BF16=0 to skip."""
    return os.environ.get("FCS_BF16", "1") != "0"


def _devices(device=None):
    """GPUs a database may use: FCS_DEVICES ("0,1,2") wins; an explicit `cuda:N` pins it to that GPU; plain `cuda`
    leaves the choice to the engine (one GPU for small databases, all of them at TED scale)."""
    env = os.environ.get("FCS_DEVICES")
    if env:
        return [int(x) for x in env.split(",") if x.strip() != ""]
    d = str(device) if device is not None else ""
    if d.startswith("cuda:"):
        return [int(d.split(":", 1)[1])]
    return None


class ResidentDatabase:
    """A database resident in HBM (all visible GPUs), standing in for the reference's
    ``target_dict['database']`` tensor: it answers ``size(0)``, ``shape`` and ``len()``."""

    def __init__(self, engine: LocalEngine, flavour: str):
        self.engine = engine
        self.flavour = flavour  # "pt" (cosine + coverage mask) or "faiss" (inner product)
        self.stamp: Optional[Tuple[int, int]] = None  # (size, mtime_ns) of the file it was loaded from

    def size(self, dim: Optional[int] = None):
        shape = (self.engine.n_rows, DIM)
        return shape if dim is None else shape[dim]

    @property
    def shape(self):
        return (self.engine.n_rows, DIM)

    def __len__(self):
        return self.engine.n_rows


def _is_cuda(device) -> bool:
    return str(device).startswith("cuda")


def read_database(db_name: str, device):
    """reference dbsearch.py:48-72.  ``.pt`` flavour: the raw [N,128] matrix is uploaded ONCE, normalised
    on the device (cosine semantics) and kept resident together with the domain lengths; the returned
    dict has the reference's keys.  ``.json`` flavour: returns the path (loaded on first search, cached)."""
    import torch

    if os.path.exists(db_name + ".pt"):
        if not _is_cuda(device):
            raise native.FcsError(native.ERR_UNSUPPORTED, f"merizo_search_b200.read_database: device {device!r} is not a CUDA "
                                                          "device and this path has no CPU implementation (use the reference's own read_database)")
        pt_path = os.path.abspath(db_name + ".pt")
        key = (pt_path, "pt")
        # <db>.index: the reference's pickle, or its one-time converted flat layout when present (dbindex.py)
        target_index, lengths = dbindex.load_index(db_name)
        resident = _resident_get(key, pt_path)
        if resident is None:
            remote = serve.connect(pt_path, "pt")  # FCS_SERVER: a resident copy in another process (serve.py)
            if remote is not None:
                resident = ResidentDatabase(remote, "pt")
                _resident_put(key, resident, pt_path)
        if resident is None:
            target_db = torch.load(db_name + ".pt", map_location="cpu")
            rows = target_db.detach().to(torch.float32).contiguous().numpy()
            assert len(target_index) == rows.shape[0]
            eng = LocalEngine(rows.shape[0], devices=_devices(device), normalise_rows=True, keep_bf16=False, has_lengths=True)
            eng.upload(0, rows, lengths)
            eng.finalize()
            resident = ResidentDatabase(eng, "pt")
            _resident_put(key, resident, pt_path)
        assert len(target_index) == resident.size(0)
        mdfn = db_name + ".metadata"
        mifn = mdfn + ".index"
        if not os.path.exists(mdfn) or not os.path.exists(mifn):
            mdfn = mifn = None
        return {"database": resident, "index": target_index, "lengths": torch.from_numpy(lengths.astype(np.float32)),
                "faiss": False, "mdfn": mdfn, "mifn": mifn}
    elif os.path.exists(db_name + ".json"):
        return {"database": db_name + ".json", "faiss": True}
    else:
        logger.error("%s is not a valid db or the path basename is incorrect; neither %s nor %s were found.",
                     db_name, db_name + ".pt", db_name + ".json")
        sys.exit(1)


def search_query_against_db(query_dict, target_dict, mincov, topk, score_corrections=None):
    """reference dbsearch.py:75-81: cosine(db, q) * (len(q.seq) >= lengths*mincov) -> topk.
    Returns {'scores': f32 tensor [k] descending, 'indices': int64 tensor [k]} (CPU tensors)."""
    import torch

    resident = target_dict["database"]
    if not isinstance(resident, ResidentDatabase):
        raise TypeError("target_dict['database'] was not produced by merizo_search_b200.read_database")
    n = resident.size(0)
    if topk > n:
        raise RuntimeError("selected index k out of range")  # torch.topk's message
    emb = query_dict["embedding"]
    if hasattr(emb, "detach"):
        emb = emb.detach().to("cpu", torch.float32).numpy()
    q = np.ascontiguousarray(emb, dtype=np.float32).reshape(1, DIM)
    qlen = np.asarray([len(query_dict["seq"])], dtype=np.int32)
    scores, ids = resident.engine.search(q, int(topk), qlen=qlen, mincov=float(mincov), qnorm=native.QNORM_COSINE,
                                         mode=native.MODE_GEMV)
    return {"scores": torch.from_numpy(scores[0]), "indices": torch.from_numpy(ids[0])}


def load_resident(db_blocks: Iterable, n_rows: int, key: Optional[Tuple[str, str]] = None, device=None) -> ResidentDatabase:
    """Upload a block iterator (reference db_iterator, dbutil.py:33-35) once; rows are used as stored
    (the faiss-flavour file is pre-normalised: dbutil.py:28-30, dbsearch.py:275)."""
    if key is not None and key in _RESIDENT:
        return _RESIDENT[key]
    t0 = time.time()
    eng = LocalEngine(n_rows, devices=_devices(device), normalise_rows=False, keep_bf16=_want_bf16(), has_lengths=False)
    eng.upload_blocks(db_blocks, progress=lambda i0: logger.info("%d DB elements, %.3f s" % (i0, time.time() - t0)))
    eng.finalize()
    resident = ResidentDatabase(eng, "faiss")
    if key is not None:
        _RESIDENT[key] = resident
    return resident


def load_resident_file(path: str, n_rows: int, device=None) -> ResidentDatabase:
    """The faiss flavour's embedding file (headerless fp32 [DB_SIZE,128], dbutil.py:28-30) straight into HBM: every shard
    reads its own byte range with positional reads from its own loader thread (no memory map, no per-block Python loop);
    cached per process by path + size + mtime."""
    path = os.path.abspath(path)
    key = (path, "faiss")
    resident = _resident_get(key, path)
    if resident is not None:
        return resident
    if os.path.getsize(path) < n_rows * DIM * 4:
        raise native.FcsError(native.ERR_INVALID, f"{path}: {os.path.getsize(path)} bytes cannot hold {n_rows} x {DIM} fp32 rows")
    if device is not None and not _is_cuda(device):
        raise native.FcsError(native.ERR_UNSUPPORTED, f"merizo_search_b200: device {device!r} is not a CUDA device and this path has no "
                                                      "CPU implementation (use the reference's own dbsearch_faiss)")
    remote = serve.connect(path, "faiss")  # FCS_SERVER: a resident copy in another process (serve.py)
    if remote is not None:
        resident = ResidentDatabase(remote, "faiss")
        _resident_put(key, resident, path)
        return resident
    t0 = time.time()
    eng = LocalEngine(n_rows, devices=_devices(device), normalise_rows=False, keep_bf16=_want_bf16(), has_lengths=False)
    eng.upload_file(path)
    eng.finalize()
    dt = time.time() - t0
    logger.info("%d DB elements resident on %d GPU(s), %.3f s (%.2f GB/s)" % (n_rows, eng.n_shards, dt, n_rows * DIM * 4 / 1e9 / max(dt, 1e-9)))
    resident = ResidentDatabase(eng, "faiss")
    _resident_put(key, resident, path)
    return resident


def knn_exact(xq, db_iterator, k, metric_type="IP", device=None, n_rows: Optional[int] = None,
              cache_key: Optional[Tuple[str, str]] = None):
    """reference knn_exact_faiss (dbsearch.py:213-248): exact inner-product kNN of xq [nq,128] against the
    rows the iterator yields (or a ResidentDatabase).  Returns (D f32 [nq,k], I int64 [nq,k]) as numpy
    arrays, each row sorted by descending inner product, ids global, (-inf,-1) padding when k > rows."""
    if metric_type not in ("IP", 0):
        logger.error("Invalid/unsupported search type: %s\n\tOnly 'IP' is currently supported." % str(metric_type))
        sys.exit(1)
    if hasattr(xq, "detach"):
        xq = xq.detach().cpu().numpy()
    xq = np.ascontiguousarray(xq, dtype=np.float32)
    logger.info("knn_exact queries size %s k=%d" % (xq.shape, k))
    t0 = time.time()
    if isinstance(db_iterator, ResidentDatabase):
        resident = db_iterator
    else:
        if n_rows is None:
            blocks = [np.asarray(b) for b in db_iterator]
            n_rows = int(sum(b.shape[0] for b in blocks))
            db_iterator = iter(blocks)
        resident = load_resident(db_iterator, n_rows, cache_key)
    D, I = resident.engine.search(xq, int(k), qnorm=native.QNORM_NONE, mode=native.MODE_AUTO)
    logger.info("kNN time: %.3f s (%d vectors)" % (time.time() - t0, resident.size(0)))
    return D, I


def release_all() -> None:
    """Free every resident database (HBM)."""
    for r in list(_RESIDENT.values()):
        r.engine.close()
    _RESIDENT.clear()


def install(reference_module=None):
    """Splice the CUDA path into the reference: ``install(programs.Foldclass.dbsearch)`` replaces its
    ``read_database`` / ``search_query_against_db`` / ``dbsearch_faiss`` module attributes (the reference's
    ``knn_exact_faiss`` is a closure and ``import faiss`` is unconditional at dbsearch.py:210, so the faiss
    flavour needs the whole driver replaced).  Returns the patched module."""
    if reference_module is None:
        import importlib

        reference_module = importlib.import_module("programs.Foldclass.dbsearch")
    from . import faiss_driver

    originals = {name: getattr(reference_module, name, None)
                 for name in ("read_database", "search_query_against_db", "dbsearch_faiss", "network_setup")}
    reference_module.read_database = read_database
    reference_module.search_query_against_db = search_query_against_db
    reference_module.dbsearch_faiss = faiss_driver.dbsearch_faiss
    # the step before the search: network_setup (dbsearch.py:35-45) hands back the batched CUDA embedder on CUDA devices
    ref_setup = originals["network_setup"]
    if callable(ref_setup) and not hasattr(ref_setup, "__wrapped__"):
        from .embed import wrap_network_setup

        reference_module.network_setup = wrap_network_setup(ref_setup)
    # Modules that did `from .dbsearch import *` BEFORE install() ran hold their own references to the originals
    # (dbsearch_fulllength.py:29; merizo.py:15 imports it at start-up): multi_domain_search (dbsearch_fulllength.py:303)
    # would keep calling the reference's read_database and put a second copy of the matrix on the GPU.  Re-bind every
    # already-imported module of the reference package whose attribute IS one of the originals.
    mod_full_name = getattr(reference_module, "__name__", None)
    if not isinstance(mod_full_name, str) or "." not in mod_full_name:
        return reference_module
    pkg = mod_full_name.rsplit(".", 1)[0]
    for mod_name, mod in list(sys.modules.items()):
        if mod is None or mod is reference_module or not (mod_name == pkg or mod_name.startswith(pkg + ".")):
            continue
        for name, orig in originals.items():
            if orig is not None and getattr(mod, name, None) is orig:
                setattr(mod, name, getattr(reference_module, name))
    return reference_module
