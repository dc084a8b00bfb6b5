"""Build recipe for libfcsearch.so (the C-ABI CUDA library), in-tree, sm_100a only.

    python -m merizo_search_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU.  The .so lands next to this file (git-ignored, but it
travels to the GPU box with the repository snapshot).  cudart is linked statically so the
library loads in any process (torch's own cudart shares the primary context with it).
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
INCLUDE = os.path.join(os.path.dirname(PKG_DIR), "include")
LIB_PATH = os.path.join(PKG_DIR, "libfcsearch.so")
STAMP_PATH = os.path.join(PKG_DIR, ".libfcsearch.stamp")

SOURCES = ["fcs_api.cu", "fcs_group.cu", "fcs_gemv.cu", "fcs_loader.cu", "fcs_merge.cu", "fcs_tc.cu", "fcs_embed.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-Xcompiler", "-pthread",
    "-cudart", "static",
] + (["-DFCS_TC_TRACE"] if os.environ.get("FCS_TC_TRACE") else []) + \
    ([f"-DFCS_TC_SLOWPATH={int(os.environ['FCS_TC_SLOWPATH'])}"] if os.environ.get("FCS_TC_SLOWPATH") else []) + \
    [f"-D{k}={int(v)}" for k, v in sorted(os.environ.items()) if k.startswith("FCS_TC_") and k in ("FCS_TC_BPOLICY",)]
# experiments: FCS_LIB_VARIANT=name builds/loads libfcsearch_<name>.so next to the default library
_VARIANT = os.environ.get("FCS_LIB_VARIANT", "")
if _VARIANT:
    LIB_PATH = os.path.join(PKG_DIR, f"libfcsearch_{_VARIANT}.so")
    STAMP_PATH = os.path.join(PKG_DIR, f".libfcsearch_{_VARIANT}.stamp")


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def _source_hash() -> str:
    h = hashlib.sha256()
    files = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))] + [os.path.join(INCLUDE, f) for f in sorted(os.listdir(INCLUDE))]
    for f in files:
        with open(f, "rb") as fh:
            h.update(os.path.basename(f).encode())  # not the absolute path: the tree is copied to other machines
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_fresh() -> bool:
    if not (os.path.exists(LIB_PATH) and os.path.exists(STAMP_PATH)):
        return False
    with open(STAMP_PATH) as fh:
        return fh.read().strip() == _source_hash()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu under csrc/ for sm_100a and link libfcsearch.so.  Returns its path."""
    if not force and is_fresh():
        return LIB_PATH
    # several ranks of one job may get here at the same time: one builds, the others wait and re-check
    import fcntl

    lock = open(os.path.join(PKG_DIR, ".build.lock"), "w")
    fcntl.flock(lock, fcntl.LOCK_EX)
    try:
        if not force and is_fresh():
            return LIB_PATH
        return _build_locked(verbose)
    finally:
        fcntl.flock(lock, fcntl.LOCK_UN)
        lock.close()


def _build_locked(verbose: bool) -> str:
    nvcc = _nvcc()
    objs = []
    build_dir = os.path.join(PKG_DIR, "build" + ("_" + _VARIANT if _VARIANT else ""))
    os.makedirs(build_dir, exist_ok=True)
    procs = []
    for src in SOURCES:
        obj = os.path.join(build_dir, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-I", INCLUDE, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = []
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            failed.append(src)
    if failed:
        raise RuntimeError(f"nvcc failed for: {', '.join(failed)}")
    link = [nvcc, "-shared", "-Xcompiler", "-pthread", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH, *objs]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("link of libfcsearch.so failed")
    with open(STAMP_PATH, "w") as fh:
        fh.write(_source_hash())
    return LIB_PATH


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
