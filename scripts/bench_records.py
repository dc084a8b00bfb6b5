"""CPU micro-benchmark of the step AFTER the search: fetching names / (start,end) / metadata records
of nq x k hits from the memory-mapped record files of a `.json` database.

  ours       merizo_search_b200.faiss_driver.RecordFiles (one fancy-indexed np.memmap read per table)
  per-hit    the reference's access pattern restated: one mmap.seek + read per hit and table
             (dbutil.py:74-103 retrieve_mmdata_by_idx, dbsearch.py:342-387); when /root/reference is importable
             (build container) the reference's own functions are timed as well.

    python scripts/bench_records.py [n_db] [n_hits] > profiles/r02_records.json
"""
import json
import mmap
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from merizo_search_b200 import faiss_driver  # noqa: E402


def build(tmp, n):
    rng = np.random.default_rng(1)
    names = [f"AF-P{i:08d}-F1-model_v4_TED{i % 7 + 1:02d}" for i in range(n)]
    with open(os.path.join(tmp, "t.index_names"), "wb") as fh:
        fh.write(b"".join(nm.ljust(32).encode() + b"\n" for nm in names))
    lens = rng.integers(25, 684, size=n)
    for stem, width in (("seq", 1), ("ca", 12)):
        end = np.cumsum(lens * width)
        idx = np.stack([end - lens * width, end], axis=1).astype(np.int64)
        idx.tofile(os.path.join(tmp, f"t_{stem}.index"))
        with open(os.path.join(tmp, f"t_{stem}.db"), "wb") as fh:
            fh.truncate(int(end[-1]))  # sparse payload: the benchmark reads offsets, names and metadata
    meta = [json.dumps({"cath": f"1.10.{i % 999}.{i % 17}"}).encode() for i in range(n)]
    mlen = np.array([len(m) for m in meta])
    mend = np.cumsum(mlen)
    np.stack([mend - mlen, mend], axis=1).astype(np.int64).tofile(os.path.join(tmp, "t_metadata.index"))
    with open(os.path.join(tmp, "t_metadata.db"), "wb") as fh:
        fh.write(b"".join(meta))
    info = {"dbfname_IP": "t.db", "DB_SIZE": n, "DB_DIM": 128, "db_names_f": "t.index_names", "sif": "t_seq.index", "sdf": "t_seq.db",
            "cif": "t_ca.index", "cdf": "t_ca.db", "mif": "t_metadata.index", "mdf": "t_metadata.db"}
    return info, lens


def per_hit(tmp, info, ids):
    """one seek + read per hit and table, like dbsearch.py:342-387 via dbutil.retrieve_*"""
    out = {}
    with open(os.path.join(tmp, info["db_names_f"]), "rb") as f:
        mm = mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ)
        names = []
        for i in ids:
            mm.seek(33 * int(i))
            names.append(mm.read(33).decode().rstrip())
        out["names"] = names
    with open(os.path.join(tmp, info["mif"]), "rb") as f, open(os.path.join(tmp, info["mdf"]), "rb") as g:
        mi = mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ)
        md = mmap.mmap(g.fileno(), 0, access=mmap.ACCESS_READ)
        metas = []
        for i in ids:
            mi.seek(16 * int(i))
            s, e = np.frombuffer(mi.read(16), dtype="int64")
            md.seek(int(s))
            metas.append(md.read(int(e - s)).decode("ascii"))
        out["metadata"] = metas
    with open(os.path.join(tmp, info["sif"]), "rb") as f:
        si = mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ)
        lens = []
        for i in ids:
            si.seek(16 * int(i))
            s, e = np.frombuffer(si.read(16), dtype="int64")
            lens.append(int(e - s))
        out["lengths"] = lens
    return out


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
    n_hits = int(sys.argv[2]) if len(sys.argv) > 2 else 65_536 * 50  # cfg5: 65,536 queries x k=50
    with tempfile.TemporaryDirectory() as tmp:
        info, lens = build(tmp, n)
        ids = np.random.default_rng(2).integers(0, n, size=n_hits)
        t0 = time.perf_counter()
        rec = faiss_driver.RecordFiles(tmp, info)
        ours = {"names": rec.names(ids), "metadata": rec.metadata(ids), "lengths": rec.lengths(ids).tolist()}
        t_ours = time.perf_counter() - t0
        sub = ids[: max(1, n_hits // 16)]  # the per-hit loop is timed on 1/16 of the hits and scaled
        t0 = time.perf_counter()
        ref = per_hit(tmp, info, sub)
        t_ref = (time.perf_counter() - t0) * (n_hits / len(sub))
        same = all(ours[k][: len(sub)] == ref[k] for k in ref)
        print(json.dumps({"db_rows": n, "hits": n_hits, "tables": ["names (33 B records)", "metadata (start,end + payload)", "lengths (start,end)"],
                          "vectorised_s": round(t_ours, 3), "per_hit_loop_s_extrapolated": round(t_ref, 3), "speedup": round(t_ref / t_ours, 1),
                          "hits_per_s_vectorised": round(n_hits / t_ours), "identical": bool(same), "cores": 1,
                          "note": "CPU only, files in the page cache; per-hit loop timed on 1/16 of the hits"}))


if __name__ == "__main__":
    main()
