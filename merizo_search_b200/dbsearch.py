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

from . import dbindex, native
from .engine import LocalEngine

logger = logging.getLogger(__name__)

DIM = native.DIM
_RESIDENT: Dict[Tuple[str, str], "ResidentDatabase"] = {}  # (abs path, flavour) -> loaded database


def _want_bf16() -> bool:
    """The bf16 operand images for the tcgen05 path cost +50 % HBM; on by default, FCS_BF16=0 to skip."""
    return os.environ.get("FCS_BF16", "1") != "0"


def _devices():
    env = os.environ.get("FCS_DEVICES")
    if env:
        return [int(x) for x in env.split(",") if x.strip() != ""]
    return None


class ResidentDatabase:
    """A database resident in HBM (all visible GPUs), standing in for the reference's
    ``target_dict['database']`` tensor: it answers ``size(0)``, ``shape`` and ``len()``."""

    def __init__(self, engine: LocalEngine, flavour: str):
        self.engine = engine
        self.flavour = flavour  # "pt" (cosine + coverage mask) or "faiss" (inner product)

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
        key = (os.path.abspath(db_name + ".pt"), "pt")
        # <db>.index: the reference's pickle, or its one-time converted flat layout when present (dbindex.py)
        target_index, lengths = dbindex.load_index(db_name)
        if key in _RESIDENT:
            resident = _RESIDENT[key]
        else:
            target_db = torch.load(db_name + ".pt", map_location="cpu")
            rows = target_db.detach().to(torch.float32).contiguous().numpy()
            assert len(target_index) == rows.shape[0]
            eng = LocalEngine(rows.shape[0], devices=_devices(), normalise_rows=True, keep_bf16=False, has_lengths=True)
            eng.upload(0, rows, lengths)
            eng.finalize()
            resident = ResidentDatabase(eng, "pt")
            _RESIDENT[key] = resident
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


def load_resident(db_blocks: Iterable, n_rows: int, key: Optional[Tuple[str, str]] = None) -> ResidentDatabase:
    """Upload a block iterator (reference db_iterator, dbutil.py:33-35) once; rows are used as stored
    (the faiss-flavour file is pre-normalised: dbutil.py:28-30, dbsearch.py:275)."""
    if key is not None and key in _RESIDENT:
        return _RESIDENT[key]
    t0 = time.time()
    eng = LocalEngine(n_rows, devices=_devices(), normalise_rows=False, keep_bf16=_want_bf16(), has_lengths=False)
    eng.upload_blocks(db_blocks, progress=lambda i0: logger.info("%d DB elements, %.3f s" % (i0, time.time() - t0)))
    eng.finalize()
    resident = ResidentDatabase(eng, "faiss")
    if key is not None:
        _RESIDENT[key] = resident
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

    reference_module.read_database = read_database
    reference_module.search_query_against_db = search_query_against_db
    reference_module.dbsearch_faiss = faiss_driver.dbsearch_faiss
    # the step before the search: network_setup (dbsearch.py:35-45) hands back the batched CUDA embedder on CUDA devices
    ref_setup = getattr(reference_module, "network_setup", None)
    if callable(ref_setup) and not hasattr(ref_setup, "__wrapped__"):
        from .embed import wrap_network_setup

        reference_module.network_setup = wrap_network_setup(ref_setup)
    return reference_module
