"""GPU: validate the AUTO cost model (fcs_api.cu api_auto_prefers_tc) against measurements: for a grid of shard sizes and
batch sizes, device time of the exact scan (FCS_MODE_GEMV) and of the tensor-core path (FCS_MODE_TC), and what AUTO picks.

    python scripts/auto_model_check.py > gpurun_out/r02_auto_model.json
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from merizo_search_b200 import native, synth  # noqa: E402


def timed(h, q, nq, k, mode, st, reps=5):
    dev = q.device
    sc = torch.empty((nq, k), dtype=torch.float32, device=dev)
    ids = torch.empty((nq, k), dtype=torch.int64, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(st):
        for _ in range(2):
            h.search_device(q.data_ptr(), nq, k, sc.data_ptr(), ids.data_ptr(), mode=mode, stream=st.cuda_stream)
        e0.record()
        for _ in range(reps):
            h.search_device(q.data_ptr(), nq, k, sc.data_ptr(), ids.data_ptr(), mode=mode, stream=st.cuda_stream)
        e1.record()
    st.synchronize()
    return e0.elapsed_time(e1) / reps, int(h.timing().last_mode)


def main():
    dev = torch.device("cuda:0")
    st = torch.cuda.Stream(dev)
    k = 10
    out = []
    for rows in (100_000, 500_000, 2_000_000, 10_000_000, 40_000_000):
        h = native.Database(rows, keep_bf16=True)
        blk = 1 << 20
        for b0 in range(0, rows, blk):
            nb = min(blk, rows - b0)
            x = synth.device_block(b0 // blk, nb, dev, base_seed=1000)
            h.upload_device(b0, nb, x.data_ptr())
            del x
        h.finalize()
        for nq in (8, 16, 32, 48, 64, 96, 128, 256, 512, 1024):
            q = torch.nn.functional.normalize(torch.randn((nq, 128), device=dev, generator=torch.Generator(dev).manual_seed(nq)))
            t_gemv, _ = timed(h, q, nq, k, native.MODE_GEMV, st, reps=3 if rows * nq > 2e9 else 5)
            t_tc, _ = timed(h, q, nq, k, native.MODE_TC, st) if nq >= 32 else (float("inf"), 0)
            _, picked = timed(h, q, nq, k, native.MODE_AUTO, st, reps=1)
            best = native.MODE_TC if t_tc < t_gemv else native.MODE_GEMV
            out.append({"rows": rows, "nq": nq, "gemv_ms": round(t_gemv, 4), "tc_ms": None if nq < 32 else round(t_tc, 4),
                        "auto_picks": "tc" if picked == native.MODE_TC else "gemv", "faster": "tc" if best == native.MODE_TC else "gemv",
                        "auto_ok": picked == best, "loss_pct": round(100 * ((t_tc if picked == native.MODE_TC else t_gemv) / min(t_tc, t_gemv) - 1), 1)})
        h.close()
        torch.cuda.empty_cache()
    wrong = [r for r in out if not r["auto_ok"]]
    print(json.dumps({"k": k, "grid": out, "wrong_choices": len(wrong), "worst_loss_pct": max([r["loss_pct"] for r in out] or [0])}))


if __name__ == "__main__":
    main()
