"""GPU (>= 2 devices): row-sharded search -- single-process LocalEngine and the torchrun/NCCL DistributedEngine."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from merizo_search_b200 import engine, native, synth
from oracle import foldclass_oracle as orc

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _need_two():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")


@pytest.mark.parametrize("n,nq,k,mode", [(50001, 3, 10, native.MODE_GEMV), (120000, 200, 20, native.MODE_TC)])
def test_local_engine_two_shards_equals_oracle(n, nq, k, mode):
    _need_two()
    db = synth.host_db(n, base_seed=81)
    q = synth.host_queries(nq, 81, normalise=True)
    eng = engine.LocalEngine(n, devices=[0, 1], keep_bf16=(mode == native.MODE_TC))
    assert eng.n_shards == 2 and eng.ranges == engine.shard_ranges(n, 2)
    eng.upload_blocks(orc.db_iterator(db, 7777))
    eng.finalize()
    s, i = eng.search(q, k, mode=mode)
    D, I = orc.knn_exact_blockwise(q, orc.db_iterator(db, 262144), k)
    full = orc.all_scores_ip(q, db)
    for r in range(nq):
        orc.check_topk(s[r], i[r], D[r], I[r], full[r], tol=1e-5)
    eng.close()


def test_torchrun_two_ranks_nccl_allgather_merge(tmp_path):
    _need_two()
    out = tmp_path / "res.npz"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29617", os.path.join(ROOT, "tests", "dist_worker.py"), str(out)]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    z = np.load(out)
    n, nq, k = int(z["n"]), int(z["nq"]), int(z["k"])
    db = synth.host_db(n, base_seed=91)
    q = synth.host_queries(nq, 91, normalise=True)
    D, I = orc.knn_exact_blockwise(q, orc.db_iterator(db, 262144), k)
    full = orc.all_scores_ip(q, db)
    for tag in ("gemv1", "tc1", "gemv2", "tc2"):  # row-sharded (1 query group) and replicated (2 query groups)
        for r_ in range(nq):
            orc.check_topk(z[f"s_{tag}"][r_], z[f"i_{tag}"][r_], D[r_], I[r_], full[r_], tol=1e-5)
