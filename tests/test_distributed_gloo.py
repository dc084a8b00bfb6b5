"""CPU, world_size 2, gloo: the multi-rank plumbing of DistributedEngine (partition, id offsets, key
all-gather layout, merge) with the oracle standing in for the per-rank CUDA search."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from merizo_search_b200 import engine, synth
from host_merge import merge_keys_host
from oracle import foldclass_oracle as orc

N, NQ, K = 5001, 7, 9


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir, q_groups):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        db = synth.host_db(N, base_seed=12)
        q = torch.from_numpy(synth.host_queries(NQ, 12, normalise=True))
        eng = engine.DistributedEngine(N, create_handle=False, query_groups=q_groups)
        lo, hi = eng.row0, eng.row1
        assert (lo, hi) == engine.shard_ranges(N, world // q_groups)[rank // q_groups]

        def local_search(qt, nq, k):  # oracle on this rank's rows, global ids = local + offset
            D, I = orc.knn_exact_blockwise(qt.numpy(), orc.db_iterator(db[lo:hi], 1024), k)
            return torch.from_numpy(engine.encode_keys(D, np.where(I >= 0, I + lo, -1)).view(np.int64))

        def merge(gathered, k):
            return merge_keys_host(gathered.numpy().view(np.uint64), k)

        s, i = eng.search(q, K, local_search=local_search, merge=merge)
        np.save(os.path.join(out_dir, f"s{rank}.npy"), s)
        np.save(os.path.join(out_dir, f"i{rank}.npy"), i)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("q_groups", [1, 2])  # 2 row shards x 1 query group, and 1 shard replicated x 2 query groups
def test_two_rank_search_equals_single_shard_oracle(tmp_path, q_groups):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path), q_groups), nprocs=world, join=True)
    db = synth.host_db(N, base_seed=12)
    q = synth.host_queries(NQ, 12, normalise=True)
    D, I = orc.knn_exact_blockwise(q, orc.db_iterator(db, 262144), K)
    full = orc.all_scores_ip(q, db)
    for rank in range(world):
        s = np.load(tmp_path / f"s{rank}.npy")
        i = np.load(tmp_path / f"i{rank}.npy")
        for r in range(NQ):
            orc.check_topk(s[r], i[r], D[r], I[r], full[r], tol=1e-6)


# ---------------------------------------------------------------------------------------------- embedding step
def test_embed_partition_is_contiguous_and_cost_balanced():
    lens = [10] * 90 + [300] * 10
    parts = engine.embed_partition(lens, 4)
    assert parts[0][0] == 0 and parts[-1][1] == len(lens) and all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
    cost = [sum(l * l for l in lens[lo:hi]) for lo, hi in parts]
    assert max(cost) <= 0.5 * sum(cost), cost                   # the ten long chains are spread over the ranks
    assert engine.embed_partition([5, 5], 4)[-1][1] == 2        # more ranks than structures: trailing slices may be empty
    assert engine.embed_partition([], 2) == [(0, 0), (0, 0)]


def _embed_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import foldclass_embed_oracle as eorc

        sd = synth.synthetic_state_dict(2024)
        chains = synth.synthetic_chains([40, 9, 33, 17, 64, 5, 21], seed=8)
        calls = []

        def embed_fn(cs):  # the oracle stands in for the per-rank CUDA embedder
            calls.append(len(cs))
            return eorc.forward_batch(cs, sd, factored=True)

        out = engine.distributed_embed(chains, embed_fn)
        assert calls and calls[0] < len(chains), "every rank embedded the whole batch"
        np.save(os.path.join(out_dir, f"e{rank}.npy"), out.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_embedding_equals_single_process_oracle(tmp_path):
    from oracle import foldclass_embed_oracle as eorc

    world, port = 2, _free_port()
    mp.spawn(_embed_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    sd = synth.synthetic_state_dict(2024)
    chains = synth.synthetic_chains([40, 9, 33, 17, 64, 5, 21], seed=8)
    want = eorc.forward_batch(chains, sd, factored=True)
    e0, e1 = np.load(tmp_path / "e0.npy"), np.load(tmp_path / "e1.npy")
    assert np.array_equal(e0, e1), "ranks disagree on the gathered embeddings"
    assert np.array_equal(e0, want)
