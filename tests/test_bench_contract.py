"""CPU: the reference arm of bench.py (the only arm that runs without a GPU) prints ONE JSON line with the keys the
driver reads, for a search workload and for the embedding workload."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    env = dict(os.environ, OMP_NUM_THREADS="4")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", *args], capture_output=True, text=True,
                       timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    return json.loads(lines[0])


@pytest.mark.timeout(600)
def test_reference_arm_search_line():
    d = _run("--workload", "cfg2", "--rows", "20000", "--steps", "1", "--warmup", "3")
    assert d["impl"] == "reference" and d["unit"] == "queries/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["config"]["workload"] == "cfg2" and d["n_gpus"] == 1
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


@pytest.mark.timeout(600)
def test_reference_arm_embed_line():
    d = _run("--workload", "embed", "--nq", "4")
    assert d["impl"] == "reference" and d["unit"] == "structures/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0


def test_both_arms_print_the_same_config_keys():
    """The driver compares the `config` of the two arms: the reference line must carry the keys of the GPU line."""
    sys.path.insert(0, ROOT)
    import bench

    wl = dict(bench.WORKLOADS["cfg3"])
    qg, shards = bench.layout(wl, wl["rows"], 8, 0)
    assert (qg, shards) == (8, 1)                                   # a 7.7 GB database is replicated, the batch split
    assert bench.layout(wl, wl["rows"], 8, 1) == (1, 8)             # --query-groups 1: pure row sharding
    assert bench.layout(bench.WORKLOADS["cfg4"], 365_000_000, 8, 0) == (1, 8)  # TED-scale slices are never replicated
    d = _run("--workload", "cfg3", "--rows", "300000", "--steps", "1", "--warmup", "3")
    wl["rows"] = 300000
    gpu_cfg = bench.make_config(wl, "cfg3", 300000, 300000, 1, 1, 1, wl["nq"])  # what run_gpu() puts into its line at N=1
    assert d["config"] == gpu_cfg                                   # identical, key for key and value for value
    assert d["config"]["workload"] == "cfg3" and d["config"]["rows_total"] == 300000 and d["config"]["k"] == 100


def test_synthetic_database_is_one_global_block_matrix():
    """Shards take slices of GLOBAL 2^20-row blocks (block seed = base + block id), so a row's content does not depend on
    how many ranks there are -- the planted-row parity check and the brute force rely on it."""
    sys.path.insert(0, ROOT)
    import bench

    blk = 1 << 20
    rows_total = 10_000_000
    for world in (1, 2, 8):
        per = -(-rows_total // world)
        seen = []
        for r in range(world):
            r0, r1 = min(rows_total, r * per), min(rows_total, (r + 1) * per)
            for gb, nrows_b, lo, hi in bench.shard_blocks(r0, r1, rows_total, blk):
                assert 0 <= lo < hi <= nrows_b <= blk
                seen.append((gb * blk + lo, gb * blk + hi))
        seen.sort()
        assert seen[0][0] == 0 and seen[-1][1] == rows_total
        assert all(a[1] == b[0] for a, b in zip(seen, seen[1:]))
    ids = bench.planted_ids(rows_total)
    assert len(ids) == bench.PLANTED and len({i * 8 // rows_total for i in ids}) == 8  # one planted row in every eighth


def test_the_in_bench_parity_check_catches_wrong_results():
    """bench.parity_check is what lets the driver's N-GPU lines carry correctness: exercise the checker itself on CPU
    tensors -- a correct result passes, a wrong id, a wrong score and a lost planted row fail."""
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch

    import bench
    from merizo_search_b200 import synth

    dev, blk, rows_total, nq, k = torch.device("cpu"), 4096, 20_000, 32, 10
    wl = {"mask": False}
    q = torch.nn.functional.normalize(torch.randn((nq, 128), generator=torch.Generator().manual_seed(1)))
    planted = bench.plant_queries(q, rows_total, dev, synth, blk, 1000)
    assert planted == bench.planted_ids(rows_total)
    db = torch.cat([synth.device_block(gb, n, dev, base_seed=1000)[lo:hi] for gb, n, lo, hi in bench.shard_blocks(0, rows_total, rows_total, blk)])
    sc, ids = torch.topk(q @ db.T, k, dim=1)

    def check(sc_, ids_):
        return bench.parity_check(sc_, ids_, q, k, wl, rows_total, 0, rows_total, None, dev, synth, blk, 1000, 1, None, torch)

    good = check(sc, ids)
    assert good["ok"] and good["planted_queries"] == bench.PLANTED and good["brute_force_queries"] >= 4, good
    assert [int(ids[j, 0]) for j in range(bench.PLANTED)] == planted
    bad_ids = ids.clone()
    bad_ids[0, 3] = (bad_ids[0, 3] + 1) % rows_total          # a wrong row at rank 3 of a sampled query
    bad_sc = sc.clone()
    bad_sc[nq - 1, 0] += 1e-3                                  # a wrong score
    lost = ids.clone()
    lost[5, 0] = lost[5, 1]                                    # planted query 5 no longer returns its row first
    res_ids, res_sc, res_lost = check(sc, bad_ids), check(bad_sc, ids), check(sc, lost)
    assert not res_sc["ok"] and not res_lost["ok"], (res_sc, res_lost)
    assert not res_ids["ok"] and res_ids["id_mismatches_beyond_ties"] == 1, res_ids  # right score, wrong row: not a tie
