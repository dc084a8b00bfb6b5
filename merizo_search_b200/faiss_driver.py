"""``dbsearch_faiss``-compatible driver for ``.json`` (larger-than-memory) databases, without faiss.

Mirrors the call contract of the reference's ``dbsearch_faiss`` (dbsearch.py:203-207 arguments,
dbsearch.py:472 return value, Appendix B of SURVEY.md for the hit-dict keys) so that
``run_dbsearch`` can call it unchanged, but:

  * the embedding matrix is mapped and uploaded ONCE per process (row-sharded over the visible
    GPUs, cached by path) instead of being paged in on every search (dbsearch.py:233-243);
  * the kNN runs in libfcsearch (``knn_exact``); no ``import faiss``;
  * hit records (names / sequences / coordinates / metadata) are fetched with vectorised
    ``np.memmap`` reads over the same on-disk formats (dbutil.py:37-43 sketches these readers)
    instead of one ``mmap.seek/read`` per hit (dbsearch.py:342-387).

PDB I/O and TM-align stay the reference's own code: they are imported from the reference package at call
time (``programs.Foldclass.utils``), exactly as the reference does.  The query embedding runs as ONE batched
CUDA forward when ``network`` is a ``merizo_search_b200.embed.FoldClassEmbedder``.
"""
from __future__ import annotations

import json
import logging
import os
import sys
from typing import List, Optional

import numpy as np

from . import dbsearch as _dbs
from . import native

logger = logging.getLogger(__name__)


# ------------------------------------------------------------------------------- on-disk formats
def read_dbinfo(dbinfo_path: str) -> dict:
    """``<db>.json``: dbfname_IP, DB_SIZE, DB_DIM, db_names_f, sif/sdf, cif/cdf, mif/mdf (dbutil.py:24)."""
    with open(dbinfo_path, "r") as fh:
        return json.load(fh)


def embedding_memmap(filename: str, n_rows: int, dim: int) -> np.memmap:
    """Headerless fp32 C-order [DB_SIZE, DB_DIM], rows pre-normalised (dbutil.py:28-30)."""
    return np.memmap(filename, dtype=np.float32, mode="r", shape=(n_rows, dim))


def row_blocks(emb: np.ndarray, batch_size: int):
    """Consecutive row blocks, like the reference's db_iterator (dbutil.py:33-35)."""
    for i0 in range(0, emb.shape[0], batch_size):
        yield emb[i0:i0 + batch_size]


class RecordFiles:
    """Vectorised readers for the per-domain record files of a ``.json`` database."""

    def __init__(self, db_dir: str, dbinfo: dict):
        self.n = int(dbinfo["DB_SIZE"])
        self.dir = db_dir
        self.info = dbinfo
        self._names = np.memmap(os.path.join(db_dir, dbinfo["db_names_f"]), dtype="S33", mode="r", shape=(self.n,))
        self._cache = {}

    def _pair(self, index_key: str, data_key: str):
        if index_key not in self._cache:
            idx = np.memmap(os.path.join(self.dir, self.info[index_key]), dtype=np.int64, mode="r", shape=(self.n, 2))
            dat = np.memmap(os.path.join(self.dir, self.info[data_key]), dtype=np.uint8, mode="r")
            self._cache[index_key] = (idx, dat)
        return self._cache[index_key]

    def names(self, ids: np.ndarray) -> List[str]:
        """33-byte records: id left-justified, space padded, newline (dbutil.py:41-43, 107-108).  One fancy-indexed read of
        the record table, one vectorised decode + strip (no per-hit Python work)."""
        recs = np.asarray(self._names[np.asarray(ids, dtype=np.int64)])
        return np.char.rstrip(np.char.decode(recs, "ascii")).tolist()

    def _spans(self, index_key: str, data_key: str, ids: np.ndarray):
        idx, dat = self._pair(index_key, data_key)
        se = np.asarray(idx[np.asarray(ids, dtype=np.int64)])  # one fancy-indexed read for all hits
        return se[:, 0], se[:, 1], dat

    def _gather_bytes(self, index_key: str, data_key: str, ids: np.ndarray, chunk_bytes: int = 64 << 20):
        """Payload bytes of every hit, concatenated: (buffer: bytes, offsets: int64 [n+1]).  The variable-length records are
        gathered with ONE fancy-indexed read per chunk of hits (flat byte index = repeat(start) + position inside the
        record), not one seek/read per hit (dbutil.py:131-145 retrieve_bytes)."""
        starts, ends, dat = self._spans(index_key, data_key, ids)
        lens = (ends - starts).astype(np.int64)
        offs = np.zeros(len(lens) + 1, dtype=np.int64)
        np.cumsum(lens, out=offs[1:])
        parts = []
        i0 = 0
        while i0 < len(lens):
            i1 = int(np.searchsorted(offs, offs[i0] + chunk_bytes, side="right")) - 1
            i1 = max(i1, i0 + 1)
            n_b = int(offs[i1] - offs[i0])
            if n_b:
                flat = np.repeat(starts[i0:i1] - (offs[i0:i1] - offs[i0]), lens[i0:i1]) + np.arange(n_b, dtype=np.int64)
                parts.append(np.asarray(dat[flat]).tobytes())
            i0 = i1
        return b"".join(parts), offs

    def _ranges(self, index_key: str, data_key: str, ids: np.ndarray):
        buf, offs = self._gather_bytes(index_key, data_key, ids)
        return [buf[int(a):int(b)] for a, b in zip(offs[:-1], offs[1:])]

    def _strings(self, index_key: str, data_key: str, ids: np.ndarray) -> List[str]:
        buf, offs = self._gather_bytes(index_key, data_key, ids)
        text = buf.decode("ascii")  # ASCII: byte offsets are character offsets
        o = offs.tolist()
        return [text[o[i]:o[i + 1]] for i in range(len(o) - 1)]

    def lengths(self, ids) -> np.ndarray:
        """Domain lengths of the hits straight from the (start,end) table: no payload bytes are touched."""
        idx, _ = self._pair("sif", "sdf")
        se = np.asarray(idx[np.asarray(ids, dtype=np.int64)])
        return (se[:, 1] - se[:, 0]).astype(np.int64)

    def sequences(self, ids) -> List[str]:
        return self._strings("sif", "sdf", ids)

    def coords(self, ids) -> List[np.ndarray]:
        out = []
        for b in self._ranges("cif", "cdf", ids):
            d = np.frombuffer(b, dtype=np.float32)
            assert d.size % 3 == 0
            out.append(d.reshape(-1, 3))
        return out

    def has_metadata(self) -> bool:
        return "mdf" in self.info and "mif" in self.info

    def metadata(self, ids) -> List[str]:
        return self._strings("mif", "mdf", ids)


def threshold_hits(D: np.ndarray, I: np.ndarray, mincos: float):
    """Query-major flattening of the hits with score >= mincos; rank order kept (dbsearch.py:318-326)."""
    rows, cols = np.nonzero(D >= mincos)
    return I[rows, cols], D[rows, cols], rows


def _reference_utils():
    try:
        from programs.Foldclass import utils as ref_utils  # the reference package must be importable
    except Exception as exc:  # pragma: no cover - depends on the user's installation
        raise ImportError("dbsearch_faiss needs the reference's programs.Foldclass.utils (read_pdb, write_pdb, "
                          "run_tmalign) on sys.path") from exc
    return ref_utils


def embed_queries(queries, network, device, inputs_are_ca: bool, pdb_chains: List[str]):
    """Query embedding step of the reference driver (dbsearch.py:287-301).  With a ``FoldClassEmbedder`` as
    ``network`` (what ``install()`` arranges on CUDA devices) the whole batch is embedded by one call of the
    batched CUDA forward; any other ``network`` (the reference's torch module) is called once per query like
    the reference does."""
    import torch

    from .embed import FoldClassEmbedder

    ref_utils = None
    query_dicts = []
    for i, qy in enumerate(queries):
        if inputs_are_ca:
            qd = qy
        else:
            ref_utils = ref_utils or _reference_utils()
            qd = ref_utils.read_pdb(pdbfile=qy, pdb_chain=pdb_chains[i])
        query_dicts.append(qd)
    if isinstance(network, FoldClassEmbedder):
        return query_dicts, network.embed_structures([qd["coords"] for qd in query_dicts])
    emb = np.zeros((len(queries), native.DIM), dtype=np.float32)
    with torch.no_grad():
        for i, qd in enumerate(query_dicts):
            x = torch.from_numpy(qd["coords"]).unsqueeze(0).to(device)
            emb[i] = network(x).detach().to("cpu", torch.float32).numpy().reshape(-1)
    return query_dicts, emb


def dbsearch_faiss(queries: list, target_dict: dict, tmp: str, network, topk: int, mincov: float, mincos: float,
                   mintm: float, fastmode: bool, device, inputs_are_ca: bool = False, search_batchsize: int = 262144,
                   search_type="IP", pdb_chain: Optional[str] = "A", skip_tmalign=False, score_corrections=None):
    if len(queries) == 0:
        logger.error("No inputs were provided!")
        sys.exit(1)
    if not os.path.exists(tmp):
        os.mkdir(tmp)
    if search_type != "IP":
        logging.error("Invalid/unsupported faiss search type: " + str(search_type) + "\n\tOnly 'IP' is currently supported.")
        sys.exit(1)
    nq = len(queries)
    dbinfofname = target_dict["database"]
    dbinfo = read_dbinfo(dbinfofname)
    db_dir = os.path.dirname(dbinfofname)
    dbfname = os.path.join(db_dir, dbinfo["dbfname_IP"])

    if pdb_chain:
        pdb_chains = pdb_chain.rstrip(",").split(",")
        if len(pdb_chains) == 1 and nq > 1:
            pdb_chains = pdb_chains * nq
    else:
        pdb_chains = ["A"] * nq

    # resident, row-sharded database: read from the file once per process (and per version of the file) by the native
    # loader -- one reader thread per GPU, positional reads into pinned staging (dbutil.py:28-35 pages it in per search)
    if int(dbinfo["DB_DIM"]) != native.DIM:
        logger.error("DB_DIM %s is not supported (FoldClassNet embeddings are %d-d)" % (dbinfo["DB_DIM"], native.DIM))
        sys.exit(1)
    resident = _dbs.load_resident_file(dbfname, int(dbinfo["DB_SIZE"]), device)

    query_dicts, q_emb = embed_queries(queries, network, device, inputs_are_ca, pdb_chains)

    # F.normalize(query_embeddings) (dbsearch.py:303-304) is fused into the search kernels (FCS_QNORM_L2)
    D, I = resident.engine.search(q_emb, int(topk), qnorm=native.QNORM_L2, mode=native.MODE_AUTO)

    hit_indices, hit_distances, query_indices = threshold_hits(D, I, mincos)
    n_hits = len(query_indices)
    if n_hits == 0:
        return [], []

    rec = RecordFiles(db_dir, dbinfo)
    logger.info("Retrieve domain hits...")
    hit_ids = rec.names(hit_indices)
    # sequences and coordinates are only needed to write the PDB files TM-align reads; without TM-align the target lengths
    # come from the offset table alone (3.3 M hits at 65,536 queries x k=50: no per-hit payload reads)
    hit_seqs = rec.sequences(hit_indices) if not skip_tmalign else None
    hit_coords = rec.coords(hit_indices) if not skip_tmalign else None
    hit_metadata = rec.metadata(hit_indices) if rec.has_metadata() else ["{ }"] * n_hits
    hit_lengths = rec.lengths(hit_indices).tolist()

    n_q = int(query_indices.max()) + 1
    results = [dict() for _ in range(n_q)]
    all_results = [dict() for _ in range(n_q)]
    counts = [0] * n_q
    n_excluded = 0
    ref_utils = None if skip_tmalign else _reference_utils()
    if not skip_tmalign:
        logger.info("TM-align top hits...")

    def base(name):
        return os.path.basename(name).replace(".pdb", "")

    for i in range(n_hits):
        qi = int(query_indices[i])
        qd = query_dicts[qi]
        hit = {
            "query": base(qd["name"]), "target": base(hit_ids[i]), "score": hit_distances[i],
            "q_len": len(qd["seq"]), "t_len": hit_lengths[i], "tmalign_output": None,
            "dom_str": qd.get("dom_str"), "dom_conf": qd.get("dom_conf"), "dom_plddt": qd.get("dom_plddt"),
            "dbindex": hit_indices[i], "metadata": hit_metadata[i],
        }
        if skip_tmalign:
            results[qi][counts[qi]] = hit
            counts[qi] += 1
            continue
        query_fn = ref_utils.write_pdb(tmp, qd["coords"], qd["seq"], name=os.path.basename(qd["name"]))
        target_fn = ref_utils.write_pdb(tmp, hit_coords[i], hit_seqs[i], name=hit_ids[i])
        tm = ref_utils.run_tmalign(query_fn, target_fn, options="-fast" if fastmode else None, keep_pdbs=False)
        hit["tmalign_output"] = tm
        if max(tm["qtm"], tm["ttm"]) >= mintm:
            results[qi][counts[qi]] = hit
            counts[qi] += 1
        else:
            all_results[qi][n_excluded] = hit
            n_excluded += 1
    if n_excluded > 0:
        logger.info("Excluded " + str(n_excluded) + " hits (across all query domains) by TM-score threshold(>=" + str(mintm) + ")")
    return results, all_results
