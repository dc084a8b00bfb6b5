"""torchrun worker for tests/test_multigpu_gpu.py: one rank per GPU, shard upload, NCCL key all-gather, GPU merge."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from merizo_search_b200 import engine, native, synth  # noqa: E402


def main():
    out = sys.argv[1]
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n, nq, k = 90001, 130, 10
    db = synth.host_db(n, base_seed=91)  # every rank regenerates the same synthetic matrix and keeps its slice
    q = torch.from_numpy(synth.host_queries(nq, 91, normalise=True)).to(dev)
    res = {}
    for qg in (1, dist.get_world_size()):  # pure row sharding, then full replication with the queries split
        eng = engine.DistributedEngine(n, device=local, keep_bf16=True, query_groups=qg)
        eng.db.upload(0, db[eng.row0:eng.row1])
        eng.db.finalize()
        for tag, mode in (("gemv", native.MODE_GEMV), ("tc", native.MODE_TC)):
            s, i = eng.search(q, k, mode=mode)
            torch.cuda.synchronize()
            res[f"s_{tag}{qg}"] = s.cpu().numpy()
            res[f"i_{tag}{qg}"] = i.cpu().numpy()
            # the end-to-end call (pinned host buffers, one sync) must give the same answer as the device-tensor call
            sh, ih = eng.search_host(q.cpu().numpy(), k, mode=mode)
            assert np.array_equal(ih, res[f"i_{tag}{qg}"]) and np.array_equal(sh, res[f"s_{tag}{qg}"]), f"search_host differs ({tag}, Q={qg})"
        eng.db.close()
    gathered = [None] * dist.get_world_size()
    dist.all_gather_object(gathered, {k_: v.tobytes() for k_, v in res.items()})
    assert all(g == gathered[0] for g in gathered), "ranks disagree on the merged result"
    if dist.get_rank() == 0:
        np.savez(out, n=n, nq=nq, k=k, **res)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
