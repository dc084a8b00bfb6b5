"""Cross-invocation residency: a small server that keeps ONE database resident in HBM and answers
searches over a unix socket, so that successive `merizo.py search` invocations do not reload it (the reference re-reads the
`.pt` / pages the 187 GB TED matrix in on EVERY call: dbsearch.py:48-64, 233-243; dbsearch_fulllength.py:303).

    python -m merizo_search_b200.serve <db basename | db.pt | db.json> [--socket /tmp/fcs.sock] [--devices 0,1,...]

Clients: with ``FCS_SERVER=/tmp/fcs.sock`` in the environment, ``read_database`` / ``dbsearch_faiss`` (the drop-in
callables of dbsearch.py / faiss_driver.py) hand the search to the server instead of loading the matrix, provided the
server holds the same file (path, size and mtime are compared).  Index / metadata / record files stay client-side: they are
memory-mapped, not loaded.  The wire format is ``multiprocessing.connection`` (length-prefixed pickles of numpy arrays)
with an authentication key (``FCS_SERVER_KEY``, default derived from the user id); it is a local, same-user convenience,
not a network service.  All arithmetic happens in the server's LocalEngine (libfcsearch); there is no CPU path here either.
"""
from __future__ import annotations

import argparse
import json
import logging
import os
import sys
from multiprocessing.connection import Client, Listener
from typing import Optional, Tuple

import numpy as np

from . import native

logger = logging.getLogger(__name__)
DIM = native.DIM


def _authkey() -> bytes:
    return os.environ.get("FCS_SERVER_KEY", f"fcs-{os.getuid()}").encode()


def file_stamp(path: str) -> Tuple[int, int]:
    st = os.stat(path)
    return (st.st_size, st.st_mtime_ns)


# ------------------------------------------------------------------------------------------------ server
def serve(engine, socket_path: str, info: dict, ready=None) -> None:
    """Answer requests on `socket_path` with `engine.search` until a shutdown request.  `engine` is any object with the
    LocalEngine search contract; `info` (path, flavour, stamp, n_rows) is what clients check before use.  One thread per
    connection; searches are serialised (a database handle is not re-entrant)."""
    import threading

    if os.path.exists(socket_path):
        os.unlink(socket_path)
    lock = threading.Lock()
    stop = threading.Event()

    def handle(conn):
        with conn:
            try:
                while not stop.is_set():
                    req = conn.recv()
                    op = req.get("op")
                    if op == "info":
                        conn.send({"ok": True, **info})
                    elif op == "search":
                        try:
                            with lock:
                                s, i = engine.search(req["q"], int(req["k"]), qlen=req.get("qlen"), mincov=float(req.get("mincov", 0.0)),
                                                     qnorm=int(req.get("qnorm", native.QNORM_NONE)), mode=int(req.get("mode", native.MODE_AUTO)),
                                                     kprime=int(req.get("kprime", 0)))
                            conn.send({"ok": True, "scores": s, "ids": i})
                        except Exception as exc:  # the client re-raises; the server keeps serving
                            conn.send({"ok": False, "error": f"{type(exc).__name__}: {exc}"})
                    elif op == "shutdown":
                        conn.send({"ok": True})
                        stop.set()
                        try:  # wake the accept loop
                            Client(socket_path, family="AF_UNIX", authkey=_authkey()).close()
                        except Exception:
                            pass
                    else:
                        conn.send({"ok": False, "error": f"unknown op {op!r}"})
            except (EOFError, OSError):
                pass  # client went away

    with Listener(socket_path, family="AF_UNIX", authkey=_authkey()) as listener:
        if ready is not None:
            ready.set()
        while not stop.is_set():
            try:
                conn = listener.accept()
            except Exception:
                continue  # a client that failed authentication or vanished during the handshake
            if stop.is_set():
                conn.close()
                break
            threading.Thread(target=handle, args=(conn,), daemon=True).start()
    try:
        os.unlink(socket_path)
    except OSError:
        pass


def _load(db: str, devices):
    """(engine, info) for a `.pt` basename / file or a `.json` database description."""
    from . import dbindex
    from .engine import LocalEngine

    base = db[:-3] if db.endswith(".pt") else (db[:-5] if db.endswith(".json") else db)
    if os.path.exists(base + ".pt"):
        import torch

        path = os.path.abspath(base + ".pt")
        _, lengths = dbindex.load_index(base)
        rows = torch.load(path, map_location="cpu").detach().to(torch.float32).contiguous().numpy()
        eng = LocalEngine(rows.shape[0], devices=devices, normalise_rows=True, keep_bf16=False, has_lengths=True)
        eng.upload(0, rows, lengths)
        eng.finalize()
        return eng, {"path": path, "flavour": "pt", "stamp": file_stamp(path), "n_rows": int(rows.shape[0]), "shards": eng.n_shards}
    if os.path.exists(base + ".json"):
        with open(base + ".json") as fh:
            dbinfo = json.load(fh)
        path = os.path.abspath(os.path.join(os.path.dirname(base + ".json"), dbinfo["dbfname_IP"]))
        n = int(dbinfo["DB_SIZE"])
        eng = LocalEngine(n, devices=devices, normalise_rows=False, keep_bf16=os.environ.get("FCS_BF16", "1") != "0", has_lengths=False)
        eng.upload_file(path)
        eng.finalize()
        return eng, {"path": path, "flavour": "faiss", "stamp": file_stamp(path), "n_rows": n, "shards": eng.n_shards}
    raise FileNotFoundError(f"neither {base}.pt nor {base}.json exists")


# ------------------------------------------------------------------------------------------------ client
class RemoteEngine:
    """Client side: the LocalEngine search contract over the server's socket."""

    def __init__(self, socket_path: str):
        self.socket_path = socket_path
        self._conn = Client(socket_path, family="AF_UNIX", authkey=_authkey())
        self.info = self._call({"op": "info"})
        self.n_rows = int(self.info["n_rows"])
        self.n_shards = int(self.info.get("shards", 1))

    def _call(self, req: dict) -> dict:
        self._conn.send(req)
        rep = self._conn.recv()
        if not rep.get("ok"):
            raise native.FcsError(native.ERR_STATE, f"search server at {self.socket_path}: {rep.get('error')}")
        return rep

    def holds(self, path: str, flavour: str) -> bool:
        """True if the server's database is this file, unchanged since it was loaded."""
        try:
            return (self.info["path"] == os.path.abspath(path) and self.info["flavour"] == flavour
                    and tuple(self.info["stamp"]) == file_stamp(path))
        except OSError:
            return False

    def search(self, q: np.ndarray, k: int, qlen=None, mincov: float = 0.0, qnorm: int = native.QNORM_NONE,
               mode: int = native.MODE_AUTO, kprime: int = 0):
        q = np.ascontiguousarray(q, dtype=np.float32).reshape(-1, DIM)
        rep = self._call({"op": "search", "q": q, "k": int(k), "qlen": None if qlen is None else np.asarray(qlen, dtype=np.int32),
                          "mincov": float(mincov), "qnorm": int(qnorm), "mode": int(mode), "kprime": int(kprime)})
        return rep["scores"], rep["ids"]

    def shutdown_server(self) -> None:
        self._call({"op": "shutdown"})

    def close(self) -> None:
        try:
            self._conn.close()
        except Exception:
            pass

    def __len__(self) -> int:
        return self.n_rows


def connect(path: str, flavour: str) -> Optional[RemoteEngine]:
    """RemoteEngine if ``FCS_SERVER`` names a live server that holds exactly this database file, else None."""
    sock = os.environ.get("FCS_SERVER")
    if not sock or not os.path.exists(sock):
        return None
    try:
        eng = RemoteEngine(sock)
    except Exception as exc:
        logger.warning("FCS_SERVER=%s is not answering (%s): loading the database locally", sock, exc)
        return None
    if eng.holds(path, flavour):
        return eng
    logger.warning("the search server at %s holds %s, not %s: loading the database locally", sock, eng.info.get("path"), path)
    eng.close()
    return None


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description="keep a Foldclass database resident in HBM and serve searches over a unix socket")
    ap.add_argument("db", help="database basename, <db>.pt or <db>.json")
    ap.add_argument("--socket", default=os.environ.get("FCS_SERVER", "/tmp/fcs.sock"))
    ap.add_argument("--devices", default=os.environ.get("FCS_DEVICES", ""), help="comma-separated GPU ordinals (default: the engine's choice)")
    args = ap.parse_args(argv)
    logging.basicConfig(level=logging.INFO, format="%(asctime)s %(message)s")
    devices = [int(x) for x in args.devices.split(",") if x.strip()] or None
    eng, info = _load(args.db, devices)
    logger.info("resident: %s (%s flavour, %d rows, %d GPU shard(s)); serving on %s", info["path"], info["flavour"], info["n_rows"],
                info["shards"], args.socket)
    try:
        serve(eng, args.socket, info)
    finally:
        eng.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
