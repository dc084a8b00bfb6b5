"""torchrun check (>= 2 GPUs; NOT yet run on hardware -- written at the end of round 1 when the GPU budget was spent):
engine.distributed_embed over NCCL.  Every rank embeds its cost-balanced slice with the CUDA embedder, one all-gather,
and the gathered matrix must equal the single-GPU result of the same embedder bit for bit.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 scripts/dist_embed_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from merizo_search_b200 import embed as b200_embed  # noqa: E402
from merizo_search_b200 import engine, synth  # noqa: E402


def main():
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    sd = synth.synthetic_state_dict(2024)
    chains = synth.synthetic_chains(synth.host_lengths(600, seed=21), seed=3)
    emb = b200_embed.FoldClassEmbedder(sd, device=local)
    got = engine.distributed_embed(chains, emb.embed_structures_device, device=dev)
    torch.cuda.synchronize()
    want = emb.embed_structures(chains)  # the whole batch on this rank alone
    ok = np.array_equal(got.cpu().numpy(), want)
    flags = [None] * dist.get_world_size()
    dist.all_gather_object(flags, bool(ok))
    if dist.get_rank() == 0:
        print("distributed_embed over NCCL:", "OK" if all(flags) else f"MISMATCH {flags}")
    emb.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if all(flags) else 1)


if __name__ == "__main__":
    main()
