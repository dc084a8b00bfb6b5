#!/usr/bin/env python
"""bench.py -- queries/s of the Foldclass database search on N B200s (one rank per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2|cfg3|cfg4|cfg4b|embed] [--impl reference]

Workloads (BASELINE.json `configs`, synthetic unit-norm 128-d rows, random queries):
  cfg3 (default)  10 M rows, 4096-query batch, k=100      -- tcgen05 path; largest single-GPU config
  cfg2            500 k rows (CATH scale), 1 query, k=10  -- fp32 scan (GEMV) path, coverage mask on
  cfg4            45.625 M rows PER GPU (365 M / 8), 1 query, k=10 -- fp32 scan, TED-scale slice
  cfg4b           the same slice, 1024-query batch, k=10          -- tcgen05 path
  cfg5            the same slice, 65,536-query batch, k=50        -- tcgen05 path (explicit --workload cfg5 only)
  embed           the step before the search: batched FoldClassNet embedding of 2048 structures (1 GPU)
With N>1 the database of cfg2/cfg3 is row-sharded over the ranks (strong scaling); cfg4 keeps
45.625 M rows per rank (weak scaling; at N=8 it is the full 365 M-row TED database).  The per-rank
key lists are exchanged with ONE NCCL all-gather and merged on the GPU.

One JSON line on stdout (rank 0).  `value` = queries/s with queries and database resident in HBM,
timed with CUDA events over K steps (max over ranks).  `e2e` = the same through the host-buffer API
(pinned host queries in, host results out, copies inside the timed region).  `roofline` is for the
dominant kernel (tcgen05 GEMM+filter for cfg3, fp32 scan for cfg2/cfg4).  `cpu_baseline` = the CPU
oracle (a port of the reference's torch / faiss-flat arithmetic) on this box's host cores on a bounded
sample.  `--impl reference` times only that CPU arm (the reference is pure Python + torch/faiss and
cannot travel to the GPU box).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "cfg2": dict(rows=500_000, nq=1, k=10, mode="gemv", mask=True, scaling="strong",
                 desc="BASELINE configs[1]: synthetic CATH-4.3-scale DB 500k x 128 fp32, single query, k=10, fp32 scan path"),
    "cfg3": dict(rows=10_000_000, nq=4096, k=100, mode="tc", mask=False, scaling="strong",
                 desc="BASELINE configs[2]: synthetic 10M x 128 DB, 4096-query batch, k=100, tcgen05 path"),
    "cfg4": dict(rows=45_625_000, nq=1, k=10, mode="gemv", mask=False, scaling="weak",
                 desc="BASELINE configs[3] slice: 365M/8 = 45.625M x 128 fp32 rows per GPU, single query, k=10, fp32 scan path"),
    "cfg4b": dict(rows=45_625_000, nq=1024, k=10, mode="tc", mask=False, scaling="weak",
                  desc="BASELINE configs[3] slice: 365M/8 = 45.625M x 128 rows per GPU, 1024-query batch, k=10, tcgen05 path"),
    "cfg5": dict(rows=45_625_000, nq=65_536, k=50, mode="tc", mask=False, scaling="weak",
                 desc="BASELINE configs[4] slice: 365M/8 = 45.625M x 128 rows per GPU, 65,536-query proteome-wide batch, k=50, "
                      "tcgen05 path (the full TED-scale configuration at N=8)"),
}
DEFAULT_WORKLOAD = "cfg3"
# The step BEFORE the search: embedding the query structures.  Not a BASELINE config of its own --
# it rides along as extra["embed"] (1 GPU) or runs alone with --workload embed.
EMBED_WORKLOAD = dict(n=2048, len_seed=21, chain_seed=3, weight_seed=2024,
                      desc="batched FoldClassNet(128) forward: 2048 synthetic C-alpha chains, lengths drawn like TED domains "
                           "(25..683, mean 126), seeded stand-in weights; fused edge kernel (tcgen05, or fp32 with FCS_EMBED_MODE=0)")


def make_config(wl, wl_name, rows_total, n_local, world, n_shards, qg, nq_local, extra=None):
    """The `config` object of the JSON line: identical keys for the GPU arm and the reference arm."""
    cfg = {"workload": wl_name, "description": wl["desc"], "rows_total": rows_total, "rows_per_gpu": n_local,
           "nq": wl["nq"], "k": wl["k"], "path": wl["mode"], "coverage_mask": wl["mask"],
           "l2": "inputs larger than L2 (database streamed from HBM every step)",
           "parallelism": (f"{n_shards} row shard(s) x {qg} query group(s) over {world} ranks, one NCCL all-gather of "
                           "packed keys + GPU merge") if world > 1 else "single GPU",
           "queries_per_gpu": nq_local}
    if extra:
        cfg.update(extra)
    return cfg


def layout(wl, rows_total, world, query_groups):
    """(query groups, row shards) of the N-rank layout bench.py uses for a workload (engine.DistributedEngine)."""
    from merizo_search_b200.engine import DistributedEngine

    qg = query_groups
    if world == 1:
        qg = 1
    elif wl["scaling"] == "weak":
        qg = 1  # the TED-scale slices are the row-sharded configuration of the metric: never replicate them
    elif qg <= 0:
        qg = DistributedEngine.auto_query_groups(rows_total, world, bytes_per_row=514 + (256 if wl["mode"] == "tc" else 0))
    while qg > 1 and (world % qg != 0 or qg > wl["nq"]):
        qg -= 1
    return qg, world // qg


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            d = json.load(fh)
        return dict(hbm=float(d["hbm_gbs"]), tensor=float(d["bf16_tflops_sustained"]), tensor_burst=float(d["bf16_tflops"]),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tensor=1400.0, tensor_burst=1590.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, gpu_index: int):
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(gpu_index)], stdout=self.tmp, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.tmp.flush()
        self.tmp.seek(0)
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.tmp.read().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                pw.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.unlink(self.tmp.name)
        except OSError:
            pass
        if sm:
            # samples under load = those drawing at least 60 % of the highest power seen (the sampler also catches the idle
            # edges of the timed region; under the power cap the LOADED clock is the lower one)
            top_pw = max(pw)
            load = [c for c, w in zip(sm, pw) if w >= 0.6 * top_pw] or sm
            out["sm_mhz"] = float(np.median(load))
            out["sm_max_mhz"] = float(max(mx))
            out["power_w_max"] = top_pw
            out["samples_under_load"] = len(load)
        out["reasons"] = sorted(reasons)
        out["samples"] = len(sm)
        return out


# --------------------------------------------------------------------------------------------- CPU arm
def cpu_oracle_throughput(wl, n_rows_total, budget_s=12.0):
    """Queries/s of the CPU oracle (port of the reference arithmetic) on a bounded sample, all host threads.

    torch flavour (cfg2): the reference's own three lines (cosine_similarity * mask -> topk) per query on the
    full 500k DB.  faiss flavour (cfg3/cfg4): blockwise mm + topk + merge restatement on a row sample,
    extrapolated linearly in the row count (stated in `sample`).
    """
    import torch

    from merizo_search_b200 import synth
    from oracle import foldclass_oracle as orc

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    k, nq = wl["k"], wl["nq"]
    if wl["mask"]:
        n = min(n_rows_total, 500_000)
        db = torch.from_numpy(synth.host_db(n, base_seed=1, normalise=False))
        lens = torch.from_numpy(synth.host_lengths(n).astype(np.float32))
        q = torch.from_numpy(synth.host_queries(8, 5))
        orc.search_torch_flavour(db, lens, q[0], 150, 0.7, k)  # warm
        t0, done = time.perf_counter(), 0
        while time.perf_counter() - t0 < budget_s and done < 64:
            orc.search_torch_flavour(db, lens, q[done % 8], 150, 0.7, k)
            done += 1
        dt = time.perf_counter() - t0
        qps = done / dt * (n / n_rows_total)
        sample = f"{done} queries x {n} rows, reference torch arithmetic (cosine_similarity*mask->topk), {dt:.1f} s"
        return qps, cores, sample
    n = min(n_rows_total, 262_144 * (1 if nq >= 1024 else 8))
    nq_s = min(nq, 4096)
    db = synth.host_db(n, base_seed=1)
    q = synth.host_queries(nq_s, 5, normalise=True)
    try:  # faiss itself when the box has it (it is not in this image: un-vendored dependency)
        import faiss  # noqa: F401
    except Exception:
        faiss = None
    if faiss is not None:
        faiss.omp_set_num_threads(cores)

        def faiss_pass():  # the reference's knn_exact_faiss loop (dbsearch.py:213-248): IndexFlat per block + ResultHeap
            rh = faiss.ResultHeap(nq_s, k, keep_max=True)
            for i0 in range(0, n, 262_144):
                xb = db[i0:i0 + 262_144]
                index = faiss.IndexFlat(128, faiss.METRIC_INNER_PRODUCT)
                index.add(xb)
                D, I = index.search(q, k)
                I += i0
                rh.add_result(D, I)
            rh.finalize()
            return rh.D, rh.I

        faiss_pass()
        t0, reps = time.perf_counter(), 0
        while True:
            faiss_pass()
            reps += 1
            if time.perf_counter() - t0 > budget_s * 0.5 or reps >= 20:
                break
        dt = (time.perf_counter() - t0) / reps
        qps = nq_s / dt * (n / n_rows_total)
        return qps, cores, (f"{nq_s} queries x {n} rows (of {n_rows_total}), faiss-cpu {faiss.__version__} IndexFlat(IP) + ResultHeap in the "
                            f"reference's block loop (262144-row blocks), {dt:.2f} s/pass x {reps}, extrapolated linearly in rows")
    orc.knn_exact_blockwise(q[: min(nq_s, 64)], orc.db_iterator(db[:65536], 65536), k)  # warm
    t0 = time.perf_counter()
    reps = 0
    while True:
        orc.knn_exact_blockwise(q, orc.db_iterator(db, 262_144), k)
        reps += 1
        if time.perf_counter() - t0 > budget_s * 0.5 or reps >= 20:
            break
    dt = (time.perf_counter() - t0) / reps
    qps = nq_s / dt * (n / n_rows_total)
    sample = (f"{nq_s} queries x {n} rows (of {n_rows_total}), blockwise mm+topk+merge restatement of faiss IndexFlatIP "
              f"(262144-row blocks), {dt:.2f} s/pass x {reps}, extrapolated linearly in rows")
    return qps, cores, sample


def cpu_embed_throughput(structures, sd, budget_s=10.0):
    """Structures/s of the CPU oracle (numpy port of FoldClassNet.forward in the reference's literal order:
    materialised [L,L,257] edge tensor, two dense layers) on a bounded sample, all host threads via BLAS."""
    from oracle import foldclass_embed_oracle as eorc

    eorc.forward(structures[0][:32], sd, factored=False)  # warm
    t0, done, res = time.perf_counter(), 0, 0
    while time.perf_counter() - t0 < budget_s and done < len(structures):
        eorc.forward(structures[done], sd, factored=False)
        res += structures[done].shape[0]
        done += 1
    dt = time.perf_counter() - t0
    return done / dt, os.cpu_count() or 1, (f"{done} structures ({res} residues) of the same batch, numpy port of the reference "
                                            f"FoldClassNet.forward (literal order), {dt:.1f} s")


def run_embed(args, steps=3, warmup=3, cpu_baseline=True):
    """Embedding workload on this rank's GPU (1 GPU): structures/s of the batched CUDA forward."""
    import torch

    from merizo_search_b200 import embed as b200_embed
    from merizo_search_b200 import native, synth

    wl = EMBED_WORKLOAD
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    n = args.nq or wl["n"]
    lens = synth.host_lengths(n, seed=wl["len_seed"])
    structures = synth.synthetic_chains(lens, seed=wl["chain_seed"])
    sd = synth.synthetic_state_dict(wl["weight_seed"])
    emb = b200_embed.FoldClassEmbedder(sd, device=local_rank)
    mode = int(os.environ.get("FCS_EMBED_MODE", str(native.EMBED_MODE_TC3)))  # the library's default: tcgen05 (bf16 hi/lo split), 16 generator warps
    tc_mode = mode != native.EMBED_MODE_FP32                                   # 0 = fp32 FMA pipe
    emb._emb.set_mode(mode)
    coords, offsets = native.Embedder._pack(structures)
    out_dev = torch.empty((n, 128), dtype=torch.float32, device=torch.device("cuda", local_rank))
    torch.cuda.synchronize()
    for _ in range(max(3, warmup)):
        emb._emb.embed_packed_to_device(coords, offsets, out_dev.data_ptr())
    sampler = ClockSampler(local_rank)
    dev_ms, edge_ms = [], []
    for _ in range(steps):  # device time: CUDA events on the embedder's own stream, recorded inside the library
        emb._emb.embed_packed_to_device(coords, offsets, out_dev.data_ptr())
        t = emb.timing()
        dev_ms.append(t.last_ms)
        edge_ms.append(t.last_edge_ms)
    clocks = sampler.stop()
    t = emb.timing()
    t0 = time.perf_counter()
    for _ in range(steps):  # e2e: list of host arrays in, host matrix out (pack + H2D + kernels + D2H)
        emb.embed_structures(structures)
    e2e_s = (time.perf_counter() - t0) / steps
    ms = float(np.mean(dev_ms))
    pairs = int(t.last_pairs)
    flops = 2.0 * 514 * 256 * pairs * 2  # edge MLP second layer, both EGNN layers: the only O(L^2 x 514 x 256) term
    fp32_equiv = flops / (float(np.mean(edge_ms)) * 1e-3) / 1e12
    sm_mhz = clocks.get("sm_mhz") or 1965.0
    props = torch.cuda.get_device_properties(local_rank)
    if tc_mode:
        # fp32-grade accuracy on the bf16 tensor pipe costs three MMAs per product (hi.hi + lo.hi + hi.lo): the roofline
        # fraction is quoted on the ALGORITHMIC flops (2*514*256 per pair per layer); the MMA rate actually issued (3x) is
        # reported beside it
        peaks = load_peaks()
        roof = {"bound": "tensor", "achieved": fp32_equiv, "peak": peaks["tensor"], "unit": "TFLOP/s", "frac": fp32_equiv / peaks["tensor"],
                "traffic": None, "kernel": ("embed_edge_tc_kernel" if mode == native.EMBED_MODE_TC else "embed_edge_tc2_kernel") + " (both layers of one batch)",
                "algorithmic": "2*514*256 flop per (i,j) pair per layer (fp32-grade result)",
                "issued_mma_tflops": 3.0 * fp32_equiv, "issued_mma_frac": 3.0 * fp32_equiv / peaks["tensor"],
                "peak_burst": peaks["tensor_burst"], "frac_burst": fp32_equiv / peaks["tensor_burst"],
                "peak_source": peaks["source"] + ", sustained bf16"}
    else:
        peak = props.multi_processor_count * 128 * 2 * sm_mhz * 1e6 / 1e12
        roof = {"bound": "fp32", "achieved": fp32_equiv, "peak": peak, "unit": "TFLOP/s", "frac": fp32_equiv / peak, "traffic": None,
                "kernel": "embed_edge_kernel (both layers of one batch)", "algorithmic": "2*514*256 flop per (i,j) pair per layer",
                "peak_source": f"derived: {props.multi_processor_count} SMs x 128 fp32 FMA lanes x 2 x {sm_mhz:.0f} MHz "
                               "(median SM clock sampled during the timed region); no tensor-core or HBM bound applies"}
    prof = os.path.join(ROOT, "profiles", "traffic_embed.json")
    if os.path.exists(prof):
        with open(prof) as fh:
            roof["traffic"] = json.load(fh).get("dram_bytes_per_launch")
    out = {
        "metric": "structures/s (FoldClassNet embedding)", "value": n / (ms * 1e-3), "unit": "structures/s", "n_gpus": 1,
        "steps": steps, "warmup": max(3, warmup), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16 hi/lo split x3 on tcgen05, fp32 accumulate (fp32-grade)" if tc_mode else "f32",
        "data": "synthetic (C-alpha-like random walks, seeded stand-in weights)",
        "config": {"workload": "embed", "description": wl["desc"], "structures": n, "residues": int(t.last_residues),
                   "pairs_per_layer": pairs, "layers": 2, "l2": "weights (0.5 MB per layer) stay in L2 by design; per-pair HBM traffic ~0"},
        "e2e": {"value": n / e2e_s, "unit": "structures/s", "h2d_bytes_per_step": int(coords.nbytes + offsets.nbytes),
                "d2h_bytes_per_step": int(n * 512), "ms_per_step": e2e_s * 1e3,
                "api": "merizo_search_b200.embed.FoldClassEmbedder.embed_structures (fcs_embed: host coordinates in, host embeddings out)"},
        "gpu_launches": int(t.last_launches) * steps,
        "roofline": roof,
        "clocks": clocks,
    }
    if cpu_baseline:
        v, cores, sample = cpu_embed_throughput(structures, sd, budget_s=10.0)
        out["cpu_baseline"] = {"value": v, "unit": "structures/s", "cores": cores, "kind": "port", "sample": sample}
    emb.close()
    del out_dev
    torch.cuda.empty_cache()
    return out


# --------------------------------------------------------------------------------------------- loader
def run_loader(args, rows=4_000_000):
    """Host -> HBM rate of the database loader (read_dbinfo/db_memmap/db_iterator, loaded ONCE here):
    a file of headerless fp32 rows (the faiss flavour's layout, dbutil.py:28-30) through (a) the native file loader
    (positional reads into pinned staging, one reader thread per shard) and (b) the block iterator over a numpy memmap
    (the reference's db_iterator blocks of 262144 rows fed to fcs_group_upload).  2 GB; the file sits in the page cache."""
    import tempfile

    from merizo_search_b200 import engine, synth

    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    tmpdir = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else tempfile.gettempdir()
    path = os.path.join(tmpdir, f"fcs_loader_bench_{os.getpid()}.db")
    blk = 1 << 18
    try:
        with open(path, "wb") as fh:
            for b in range(0, rows, blk):
                synth.host_db(min(blk, rows - b), base_seed=77 + b // blk).tofile(fh)
        nbytes = rows * 512
        out = {"rows": rows, "bytes": nbytes, "file": f"{tmpdir} (page cache)", "shards": 1}
        for name in ("file", "memmap_blocks"):
            best = None
            for _ in range(2):
                eng = engine.LocalEngine(rows, devices=[local_rank], keep_bf16=False)
                t0 = time.perf_counter()
                if name == "file":
                    eng.upload_file(path)
                else:
                    mm = np.memmap(path, dtype=np.float32, mode="r", shape=(rows, 128))
                    eng.upload_blocks(mm[i0:i0 + 262_144] for i0 in range(0, rows, 262_144))
                eng.finalize()
                dt = time.perf_counter() - t0
                eng.close()
                best = dt if best is None else min(best, dt)
            out[name + "_gbs"] = nbytes / best / 1e9
            out[name + "_s"] = best
        out["api"] = ("file: LocalEngine.upload_file -> fcs_group_upload_file (pread into pinned staging, H2D DMA double-buffered); "
                      "memmap_blocks: LocalEngine.upload_blocks over np.memmap -> fcs_group_upload (page-cache copy into pinned staging)")
        return out
    finally:
        try:
            os.unlink(path)
        except OSError:
            pass


# --------------------------------------------------------------------------------------------- single-process engine
def run_local(args):
    """`--workload local`: the single-process deployment (what `merizo.py search -d cuda` gets): engine.LocalEngine = the
    library's shard group over ALL visible GPUs, driven by one host thread, key lists exchanged device-to-device.  Wall-clock
    latency per call through the host-buffer API for the CATH-scale and the TED-scale databases, with a brute-force check."""
    import torch

    from merizo_search_b200 import engine, native, synth

    ngpu = torch.cuda.device_count()
    blk = 1 << 20
    res = {"gpus_visible": ngpu, "cases": []}

    def build(rows_total, devices, mask, keep_bf16):
        eng = engine.LocalEngine(rows_total, devices=devices, keep_bf16=keep_bf16, has_lengths=mask)
        lens_all = synth.host_lengths(rows_total).astype(np.int32) if mask else None
        for si, (r0, r1) in enumerate(eng.ranges):
            dev = torch.device("cuda", eng.devices[si])
            sh = eng.shard(si)
            with torch.cuda.device(dev):
                ld = torch.from_numpy(lens_all[r0:r1]).to(dev) if mask else None
                b0 = r0
                while b0 < r1:  # global 2^20-row blocks (the same block seeds as every other bench workload)
                    gb = b0 // blk
                    x = synth.device_block(gb, min(blk, rows_total - gb * blk), dev, base_seed=1000)
                    lo, hi = b0 - gb * blk, min(r1, (gb + 1) * blk) - gb * blk
                    part = x[lo:hi].contiguous()
                    sh.upload_device(b0 - r0, hi - lo, part.data_ptr(), ld[b0 - r0:b0 - r0 + hi - lo].contiguous().data_ptr() if mask else None)
                    b0 += hi - lo
                    del x, part
                torch.cuda.synchronize(dev)
        eng.finalize()
        return eng, lens_all

    def brute(rows_total, q, k, mask, lens_all):
        dev = torch.device("cuda", 0)
        qd = torch.from_numpy(q).to(dev)
        best_s = torch.full((q.shape[0], k), -float("inf"), device=dev)
        best_i = torch.full((q.shape[0], k), -1, dtype=torch.int64, device=dev)
        for b0 in range(0, rows_total, blk):
            nb = min(blk, rows_total - b0)
            x = synth.device_block(b0 // blk, nb, dev, base_seed=1000)
            sc = qd @ x.T
            if mask:
                need = torch.from_numpy(lens_all[b0:b0 + nb]).to(dev).to(torch.float32) * torch.tensor(0.7, dtype=torch.float32, device=dev)
                sc = sc * (torch.tensor(150.0, device=dev) >= need).to(torch.float32)[None, :]
            ts, ti = torch.topk(sc, min(k, nb), dim=1)
            cs, ci = torch.cat([best_s, ts], 1), torch.cat([best_i, ti + b0], 1)
            best_s, pos = torch.topk(cs, k, dim=1)
            best_i = torch.gather(ci, 1, pos)
            del x, sc
        return best_s.cpu().numpy(), best_i.cpu().numpy()

    def case(name, rows_total, devices, nq, k, mask, mode, reps):
        eng, lens_all = build(rows_total, devices, mask, mode != native.MODE_GEMV)
        g = torch.Generator().manual_seed(4242)
        q = torch.nn.functional.normalize(torch.randn((nq, 128), generator=g)).numpy()
        kw = dict(qlen=np.full(nq, 150, np.int32), mincov=0.7) if mask else {}
        for _ in range(5):
            s, i = eng.search(q, k, mode=mode, **kw)
        t0 = time.perf_counter()
        for _ in range(reps):
            s, i = eng.search(q, k, mode=mode, **kw)
        dt = (time.perf_counter() - t0) / reps
        sample = sorted(set([0, nq // 2, nq - 1]))
        ws, wi = brute(rows_total, q[sample], k, mask, lens_all)
        ok = bool(np.abs(s[sample] - ws).max() <= 1e-5 and ((i[sample] == wi) | (np.abs(s[sample] - ws) <= 1e-5)).all())
        res["cases"].append({"case": name, "rows_total": rows_total, "shards": eng.n_shards, "devices": list(eng.devices), "nq": nq, "k": k,
                             "path": "tc" if mode == native.MODE_TC else "gemv", "ms_per_call": dt * 1e3, "queries_per_s": nq / dt,
                             "parity_checked": ok, "fallback_queries": eng.group.last_fallbacks()})
        eng.close()
        torch.cuda.empty_cache()

    all_dev = list(range(ngpu))
    case("cfg2 (CATH scale), engine's own choice", 500_000, None, 1, 10, True, native.MODE_GEMV, 2000)
    if ngpu > 1:
        case(f"cfg2 (CATH scale), forced over {ngpu} GPUs", 500_000, all_dev, 1, 10, True, native.MODE_GEMV, 2000)
    per = 45_625_000
    case(f"cfg4 (TED scale: 365M/8 rows per GPU x {ngpu}), 1 query", per * ngpu, all_dev, 1, 10, False, native.MODE_GEMV, 100)
    case(f"cfg4 (TED scale: 365M/8 rows per GPU x {ngpu}), 1024-query batch", per * ngpu, all_dev, 1024, 10, False, native.MODE_TC, 10)
    case("cfg3 (10M rows), 4096-query batch, engine's own choice", 10_000_000, None, 4096, 100, False, native.MODE_TC, 10)
    res["api"] = "merizo_search_b200.engine.LocalEngine.search (fcs_group_search: host queries in, host scores/ids out, one host thread)"
    return res


# --------------------------------------------------------------------------------------------- parity (outside the timed region)
PLANTED = 8  # queries replaced by database rows (spread over the whole id range, i.e. over every shard)


def shard_blocks(r0, r1, rows_total, blk):
    """The synthetic database is a GLOBAL matrix made of 2^20-row blocks (block b is generated from seed base+b, whoever
    generates it); a shard [r0, r1) takes the slices of the blocks it overlaps.  Yields (block, rows in block, lo, hi):
    rows [lo, hi) of the block are global rows [block*blk + lo, block*blk + hi)."""
    b0 = r0
    while b0 < r1:
        gb = b0 // blk
        lo, hi = b0 - gb * blk, min(r1, (gb + 1) * blk) - gb * blk
        yield gb, min(blk, rows_total - gb * blk), lo, hi
        b0 += hi - lo


def planted_ids(rows_total):
    return [int((2 * j + 1) * rows_total // (2 * PLANTED)) for j in range(PLANTED)]


def plant_queries(q_dev, rows_total, dev, synth, blk, base_seed):
    """Overwrite the first PLANTED queries with database rows (regenerated from their block seeds: every rank can do
    this for any global id), so their top hit is known: the row itself, score 1."""
    ids = planted_ids(rows_total)
    for j, gid in enumerate(ids):
        b = gid // blk
        nb = min(blk, rows_total - b * blk)
        x = synth.device_block(b, nb, dev, base_seed=base_seed)
        q_dev[j] = x[gid - b * blk]
        del x
    return ids


def parity_check(sc, ids, q_dev, k, wl, rows_total, r0, r1, lens_local, dev, synth, blk, base_seed, world, dist, torch):
    """Result of one search (all queries, merged over ranks) against (1) the planted rows and (2) an independent
    brute-force top-k (torch matmul + topk over this rank's regenerated shard, all-gathered and merged) on a sample of
    the queries.  Same rule as the tests: scores within 1e-5, ids equal except inside score ties."""
    out = {"planted_queries": PLANTED if (not wl["mask"] and q_dev.shape[0] >= PLANTED) else 0, "brute_force_queries": 0, "max_abs_score_diff": 0.0, "id_mismatches_beyond_ties": 0, "errors": []}
    nq = q_dev.shape[0]
    sc_h, ids_h = sc.cpu().numpy(), ids.cpu().numpy()
    if not wl["mask"] and nq >= PLANTED:
        for j, gid in enumerate(planted_ids(rows_total)):
            if ids_h[j, 0] != gid or abs(float(sc_h[j, 0]) - 1.0) > 1e-5:
                out["errors"].append(f"planted query {j}: expected row {gid} at rank 0 with score 1, got {int(ids_h[j, 0])} / {float(sc_h[j, 0]):.6f}")
    if not ((np.diff(sc_h, axis=1) <= 0).all()):
        out["errors"].append("scores not sorted")
    sample = sorted(set(i for i in [0, 1, nq // 2, nq - 1] + ([PLANTED, 2 * PLANTED + 1] if nq > 2 * PLANTED + 1 else []) if 0 <= i < nq))
    qs = q_dev[sample].clone()
    best_s = torch.full((len(sample), k), -float("inf"), device=dev)
    best_i = torch.full((len(sample), k), -1, dtype=torch.int64, device=dev)
    for gb, nrows_b, lo, hi in shard_blocks(r0, r1, rows_total, blk):
        x = synth.device_block(gb, nrows_b, dev, base_seed=base_seed)[lo:hi]
        g0 = gb * blk + lo  # global id of x[0]
        s = qs @ x.T
        if wl["mask"]:  # (qlen >= lengths * mincov).float(), an fp32 product (reference dbsearch.py:76)
            need = lens_local[g0 - r0:g0 - r0 + (hi - lo)].to(torch.float32) * torch.tensor(0.7, dtype=torch.float32, device=dev)
            s = s * (torch.tensor(150.0, device=dev) >= need).to(torch.float32)[None, :]
        ts, ti = torch.topk(s, min(k, hi - lo), dim=1)
        cs, ci = torch.cat([best_s, ts], 1), torch.cat([best_i, ti + g0], 1)
        best_s, pos = torch.topk(cs, k, dim=1)
        best_i = torch.gather(ci, 1, pos)
        del x, s
    if world > 1:
        gs = torch.empty((world,) + tuple(best_s.shape), device=dev)
        gi = torch.empty((world,) + tuple(best_i.shape), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(gs.view(world * len(sample), k), best_s.contiguous())
        dist.all_gather_into_tensor(gi.view(world * len(sample), k), best_i.contiguous())
        cs = gs.permute(1, 0, 2).reshape(len(sample), world * k)
        ci = gi.permute(1, 0, 2).reshape(len(sample), world * k)
        # replicated layouts gather the same shard several times: drop duplicate ids before the final top-k
        order = torch.argsort(ci, dim=1, stable=True)
        ci, cs = torch.gather(ci, 1, order), torch.gather(cs, 1, order)
        dup = torch.zeros_like(ci, dtype=torch.bool)
        dup[:, 1:] = ci[:, 1:] == ci[:, :-1]
        cs = torch.where(dup, torch.full_like(cs, -float("inf")), cs)
        best_s, pos = torch.topk(cs, k, dim=1)
        best_i = torch.gather(ci, 1, pos)
    ws, wi = best_s.cpu().numpy(), best_i.cpu().numpy()
    got_s, got_i = sc_h[sample], ids_h[sample]
    out["brute_force_queries"] = len(sample)
    diff = np.abs(got_s - ws)
    out["max_abs_score_diff"] = float(diff.max())
    if diff.max() > 1e-5:
        out["errors"].append(f"score differs from brute force by {diff.max():.3e}")
    mism = got_i != wi
    if mism.any():
        # an id mismatch is fine only inside a tie: the reported row's TRUE score (recomputed from the regenerated row) must
        # equal the reported score and tie the brute-force score at that rank within the tolerance
        bad = 0
        for r, c in zip(*np.nonzero(mism)):
            gid = int(got_i[r, c])
            ok_tie = 0 <= gid < rows_total and abs(float(got_s[r, c]) - float(ws[r, c])) <= 1e-5
            if ok_tie:
                gb = gid // blk
                row = synth.device_block(gb, min(blk, rows_total - gb * blk), dev, base_seed=base_seed)[gid - gb * blk]
                true = float(qs[r] @ row)
                if wl["mask"]:
                    L = float(synth.host_lengths(rows_total)[gid])
                    true *= 1.0 if 150.0 >= np.float32(L) * np.float32(0.7) else 0.0
                ok_tie = abs(true - float(got_s[r, c])) <= 1e-5
            bad += 0 if ok_tie else 1
        if bad or (mism.sum(axis=1) > max(2, k // 10)).any():
            out["id_mismatches_beyond_ties"] = int(bad) if bad else int(mism.sum())
            out["errors"].append(f"{int(mism.sum())} id mismatches against brute force, {bad} of them not score ties")
    out["ok"] = not out["errors"]
    return out


# --------------------------------------------------------------------------------------------- GPU arm
def run_gpu(args, wl, wl_name, steps=None, warmup=None):
    """One workload on this rank's GPU; rank 0 returns the JSON dict (others None)."""
    args = argparse.Namespace(**vars(args))
    if steps:
        args.steps = steps
    if warmup:
        args.warmup = warmup
    import torch
    import torch.distributed as dist

    from merizo_search_b200 import native, synth
    from merizo_search_b200.engine import shard_ranges

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N > 1")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)

    peaks = load_peaks()
    nq, k = wl["nq"], wl["k"]
    rows_total = wl["rows"] * (world if wl["scaling"] == "weak" else 1)
    if args.rows:
        rows_total = args.rows
    # N>1 layout: world = R row shards x Q query groups (engine.DistributedEngine); Q=1 is pure row sharding
    from merizo_search_b200.engine import DistributedEngine

    qg, n_shards = layout(wl, rows_total, world, args.query_groups)
    r0, r1 = shard_ranges(rows_total, n_shards)[rank // qg]
    n_local = r1 - r0
    nq_local = -(-nq // qg)
    mode = native.MODE_TC if wl["mode"] == "tc" else native.MODE_GEMV

    # ---- database: generated on the device block by block, uploaded into the handle, finalized
    t_load = time.perf_counter()
    h = native.Database(n_local, device=local_rank, id_offset=r0, keep_bf16=(wl["mode"] == "tc"), has_lengths=wl["mask"])
    blk = 1 << 20
    lens_all = None
    if wl["mask"]:
        lens_all = torch.from_numpy(synth.host_lengths(rows_total)[r0:r1].astype(np.int32)).to(dev)
    for gb, nrows_b, lo, hi in shard_blocks(r0, r1, rows_total, blk):
        x = synth.device_block(gb, nrows_b, dev, base_seed=1000)[lo:hi].contiguous()
        b0 = gb * blk + lo - r0  # first local row of this slice
        h.upload_device(b0, hi - lo, x.data_ptr(), lens_all[b0:b0 + hi - lo].contiguous().data_ptr() if wl["mask"] else None)
        del x
    h.finalize()
    torch.cuda.synchronize()
    t_load = time.perf_counter() - t_load

    g = torch.Generator(device=dev)
    g.manual_seed(4242)  # same queries on every rank (replicated)
    q_dev = torch.nn.functional.normalize(torch.randn((nq, 128), device=dev, generator=g))
    if not wl["mask"] and nq >= PLANTED:
        plant_queries(q_dev, rows_total, dev, synth, blk, 1000)
    q_host = q_dev.cpu().pin_memory()
    qlen = np.full(nq, 150, np.int32) if wl["mask"] else None
    mincov = 0.7 if wl["mask"] else 0.0

    stream = torch.cuda.Stream(dev)
    keys = torch.empty((nq, k), dtype=torch.int64, device=dev)
    sc = torch.empty((nq, k), dtype=torch.float32, device=dev)
    ids = torch.empty((nq, k), dtype=torch.int64, device=dev)
    launches = 0
    deng = None
    if world > 1:
        deng = DistributedEngine(rows_total, rank=rank, world_size=world, device=local_rank, create_handle=False, query_groups=qg)
        deng.db = h
        assert (deng.row0, deng.row1) == (r0, r1)

    def step_device():
        nonlocal launches
        with torch.cuda.stream(stream):
            if world == 1:
                h.search_device(q_dev.data_ptr(), nq, k, sc.data_ptr(), ids.data_ptr(), out_keys_ptr=keys.data_ptr(),
                                qlen=qlen, mincov=mincov, mode=mode, stream=stream.cuda_stream)
                launches += launches_per_search
            else:  # shard search -> ONE NCCL all-gather of packed keys -> merge kernel, on every rank
                deng.search(q_dev, k, qlen=qlen, mincov=mincov, mode=mode)
                launches += launches_per_search + 1

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    launches_per_search = 0
    for _ in range(args.warmup):
        step_device()
    barrier()
    launches_per_search = int(h.timing().last_launches)
    # ---- parity, outside the timed region: the warm-up's last result against planted rows and a brute-force sample
    with torch.cuda.stream(stream):
        if world == 1:
            chk_sc, chk_ids = sc, ids
        else:
            chk_sc, chk_ids = deng.search(q_dev, k, qlen=qlen, mincov=mincov, mode=mode)
    barrier()
    parity = parity_check(chk_sc, chk_ids, q_dev, k, wl, rows_total, r0, r1, lens_all, dev, synth, blk, 1000, world, dist, torch)
    if world > 1:
        okt = torch.tensor([1 if parity["ok"] else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)  # every rank holds the merged result: all of them must agree
        parity["ok_all_ranks"] = bool(int(okt.item()))
        parity["ok"] = parity["ok"] and parity["ok_all_ranks"]
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches = 0
    kernel_ms = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record()
    for _ in range(args.steps):  # fully asynchronous: nothing in a search waits for the host
        step_device()
    with torch.cuda.stream(stream):
        e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler else None
    if wl["mode"] == "tc":
        # dominant-kernel time: event pairs around every GEMM+filter launch inside the library (fcs_set_profiling; off in
        # the timed region because events between kernels cost their programmatic-launch overlap).  A second back-to-back
        # loop with the events on, read once for its LAST search -- a search that ran under sustained clocks (searches run
        # one at a time with host round trips in between boost higher and would flatter the roofline)
        h.set_profiling(True)
        for _ in range(max(3, min(args.steps, 10))):
            step_device()
        barrier()
        kernel_ms.append(h.timing().last_kernel_ms)
        h.set_profiling(False)
    else:
        kernel_ms = [ms_total / args.steps]  # one kernel per step, launched back to back
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    value = nq / (ms_per_step * 1e-3)
    timing = h.timing()

    # ---- e2e: host buffers through the public API (H2D + D2H inside the timed region).
    # 1 GPU: native.Database.search (fcs_search).  N GPUs: engine.DistributedEngine.search_host (pinned host
    # queries -> H2D -> shard search -> NCCL all-gather of keys -> GPU merge -> D2H) on every rank.
    e2e_out = (np.empty((nq, k), dtype=np.float32), np.empty((nq, k), dtype=np.int64))  # a batch loop reuses its result arrays

    def step_e2e():
        if world == 1:
            return h.search(q_host.numpy(), k, qlen=qlen, mincov=mincov, mode=mode, out=e2e_out)
        with torch.cuda.stream(stream):
            out_ = deng.search_host(q_host, k, qlen=qlen, mincov=mincov, mode=mode)
        return out_

    for _ in range(max(1, args.warmup // 2)):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_s = (time.perf_counter() - t0) / args.steps
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())

    out = None
    if rank == 0:
        kms = float(np.mean(kernel_ms))
        if wl["mode"] == "tc":
            flops = 2.0 * nq_local * n_local * 128  # this GPU's share: its query slice against its row shard
            achieved = flops / (kms * 1e-3) / 1e12
            roof = dict(bound="tensor", achieved=achieved, peak=peaks["tensor"], unit="TFLOP/s", frac=achieved / peaks["tensor"],
                        traffic=None, kernel="tc_gemm_filter_kernel (all rounds of one search, timed inside the sustained loop)",
                        algorithmic="2*nq*rows*128 flop per search", peak_source=peaks["source"] + ", sustained bf16 (the kernel is timed "
                        "inside a long step)", peak_burst=peaks["tensor_burst"], frac_burst=achieved / peaks["tensor_burst"],
                        step_frac=(flops / (ms_per_step * 1e-3) / 1e12) / peaks["tensor"])
        else:
            bpr = 512 + (2 if wl["mask"] else 0)
            byts = float(n_local) * bpr
            achieved = byts / (kms * 1e-3) / 1e9
            roof = dict(bound="hbm", achieved=achieved, peak=peaks["hbm"], unit="GB/s", frac=achieved / peaks["hbm"], traffic=None,
                        kernel="gemv_topk_kernel", algorithmic=f"{bpr} B per row per launch", peak_source=peaks["source"])
        # dram bytes per launch come from an ncu --set full capture of exactly this workload on ONE GPU (profiles/); a rank
        # of an N-GPU run does a different amount of work per launch, so no figure is quoted there
        prof = os.path.join(ROOT, "profiles", f"traffic_{wl_name}.json")
        if world == 1 and not args.rows and not args.nq and os.path.exists(prof):
            with open(prof) as fh:
                tj = json.load(fh)
            roof["traffic"] = tj.get("dram_bytes_per_launch")
            roof["traffic_source"] = f"profiles/traffic_{wl_name}.json ({tj.get('captured', 'ncu --set full, one search')})"
        out = {
            "metric": "queries/s (exact top-k vs 128-d DB)", "value": value, "unit": "queries/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": wl["scaling"], "vs_baseline": None,
            "dtype": "bf16 tensor-core contraction + fp32 exact rescore" if wl["mode"] == "tc" else "f32",
            "data": "synthetic (i.i.d. N(0,1) rows, L2-normalised, generated on device per 2^20-row block; random unit queries)",
            # `config` describes the workload only and is identical, key for key and value for value, on the reference arm;
            # what this run measured about itself goes into `run`
            "config": make_config(wl, wl_name, rows_total, n_local, world, n_shards, qg, nq_local),
            "run": {"db_load_s": round(t_load, 2), "tc_fallback_queries": int(timing.last_tc_fallbacks), "tc_rounds": int(timing.last_rounds)},
            "parity_checked": bool(parity["ok"]), "parity": parity,
            "e2e": {"value": nq / e2e_s, "unit": "queries/s", "h2d_bytes_per_step": int(nq * 512),
                    "d2h_bytes_per_step": int(nq * k * 12), "ms_per_step": e2e_s * 1e3,
                    "api": ("merizo_search_b200.native.Database.search (fcs_search: pinned host queries in, host scores/ids out)"
                            if world == 1 else "merizo_search_b200.engine.DistributedEngine.search_host (H2D, shard search, "
                            "NCCL all-gather of keys, GPU merge, D2H on every rank)")},
            "gpu_launches": int(launches),
            "roofline": roof,
            "clocks": clocks,
        }
    h.close()
    del q_dev, keys, sc, ids
    torch.cuda.empty_cache()
    if world > 1:
        dist.barrier()
    return out, rows_total


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=0)
    ap.add_argument("--warmup", type=int, default=0)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("FCS_BENCH_WORKLOAD", DEFAULT_WORKLOAD), choices=sorted(WORKLOADS) + ["embed", "local"])
    ap.add_argument("--rows", type=int, default=0, help="override the total row count (debugging)")
    ap.add_argument("--nq", type=int, default=0, help="override the batch size (debugging)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the ride-along workloads (cfg4, cfg2)")
    ap.add_argument("--query-groups", type=int, default=0,
                    help="N>1: ranks = row shards x query groups; 0 = auto (replicate the database as far as ~80 GB per "
                         "GPU allow and split the batch), 1 = pure row sharding")
    args = ap.parse_args()
    if args.workload == "local":  # single-process engine over all visible GPUs (not under torchrun)
        if int(os.environ.get("RANK", "0")) == 0:
            print(json.dumps({"workload": "local", **run_local(args)}))
        return 0
    if args.workload == "embed":  # the step before the search, alone (1 GPU)
        if int(os.environ.get("RANK", "0")) != 0:
            return 0
        if args.impl == "reference":
            from merizo_search_b200 import synth

            n = args.nq or EMBED_WORKLOAD["n"]
            structures = synth.synthetic_chains(synth.host_lengths(n, seed=EMBED_WORKLOAD["len_seed"]), seed=EMBED_WORKLOAD["chain_seed"])
            v, cores, sample = cpu_embed_throughput(structures, synth.synthetic_state_dict(EMBED_WORKLOAD["weight_seed"]), budget_s=20.0)
            print(json.dumps({"impl": "reference", "metric": "structures/s (FoldClassNet embedding)", "value": v, "unit": "structures/s",
                              "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True,
                              "config": {"workload": "embed", "structures": n},
                              "cpu_baseline": {"value": v, "unit": "structures/s", "cores": cores, "kind": "port", "sample": sample},
                              "e2e": {"value": v, "unit": "structures/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
            return 0
        print(json.dumps(run_embed(args, steps=args.steps or 3, warmup=args.warmup or 3, cpu_baseline=not args.no_cpu_baseline)))
        return 0
    wl = dict(WORKLOADS[args.workload])
    if args.nq:
        wl["nq"] = args.nq
    if args.steps <= 0:  # long enough (>= ~0.2 s) for nvidia-smi to sample clocks inside the timed region
        args.steps = {"cfg3": 60, "cfg2": 8000, "cfg4": 120, "cfg4b": 40, "cfg5": 3}[args.workload]
    if args.warmup <= 0:
        args.warmup = 3 if wl["mode"] == "tc" else 10
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        if rank != 0:
            return 0
        rows_total = args.rows or wl["rows"] * (world if wl["scaling"] == "weak" else 1)
        ref_qg, ref_shards = layout(wl, rows_total, world, args.query_groups)
        vals = []
        cores = sample = None
        for _ in range(max(1, min(args.steps, 3))):
            v, cores, sample = cpu_oracle_throughput(wl, rows_total, budget_s=10.0)
            vals.append(v)
        v = float(np.median(vals))
        print(json.dumps({
            "impl": "reference", "metric": "queries/s (exact top-k vs 128-d DB)", "value": v, "unit": "queries/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wl["nq"] / v,
            "higher_is_better": True, "scaling": wl["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            # same keys and values as the GPU arm's config (the layout fields describe the GPU arm this line is compared with)
            "config": make_config(wl, args.workload, rows_total, -(-rows_total // ref_shards), world, ref_shards, ref_qg,
                                  -(-wl["nq"] // ref_qg)),
            "cpu_baseline": {"value": v, "unit": "queries/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        return 0

    out, rows_total = run_gpu(args, wl, args.workload)
    # The other BASELINE configurations ride along as `extra` (same timing rules, fewer steps): cfg4 is the metric's own
    # configuration -- 365 M x 128 fp32 rows, k=10, row-sharded, NCCL key all-gather + merge -- as a 45.6 M-row-per-GPU
    # slice (the full database at N=8); cfg2 is the CATH-scale single-query case (1 GPU only).
    extra = {}
    if not args.no_extra:
        # cfg5 (65,536 queries, k=50: BASELINE configs[4]) rides along at N=8, where the slices add up to the full 365 M-row
        # TED-scale database; on fewer GPUs it runs with --workload cfg5 (1-GPU slice: profiles/)
        others = [w for w in (("cfg4", "cfg4b", "cfg2") if world == 1 else (("cfg4", "cfg4b", "cfg5") if world == 8 else ("cfg4", "cfg4b")))
                  if w != args.workload]
        for w in others:
            try:
                o, _ = run_gpu(args, WORKLOADS[w], w, steps={"cfg4": 40, "cfg4b": 8, "cfg2": 2000, "cfg3": 10, "cfg5": 3}[w],
                               warmup=3 if w == "cfg5" else 5)
                if rank == 0:
                    extra[w] = {key: o[key] for key in ("value", "unit", "ms_per_step", "scaling", "e2e", "roofline", "config", "run", "gpu_launches",
                                                        "parity_checked", "parity")}
            except Exception as exc:  # an extra must never take the primary line down
                if rank == 0:
                    extra[w] = {"error": str(exc)[:300]}
        if world == 1 and not args.nq:
            try:
                extra["loader"] = run_loader(args)
            except Exception as exc:
                extra["loader"] = {"error": str(exc)[:300]}
        if world == 1 and not args.nq:
            try:
                o = run_embed(args, steps=3, warmup=3, cpu_baseline=not args.no_cpu_baseline)
                extra["embed"] = {key: o[key] for key in ("metric", "value", "unit", "ms_per_step", "e2e", "roofline", "config",
                                                          "gpu_launches", "cpu_baseline") if key in o}
            except Exception as exc:
                extra["embed"] = {"error": str(exc)[:300]}
    if world > 1:
        import torch.distributed as dist

        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        if not args.no_cpu_baseline and world == 1:
            v, cores, sample = cpu_oracle_throughput(wl, rows_total, budget_s=12.0)
            out["cpu_baseline"] = {"value": v, "unit": "queries/s", "cores": cores, "kind": "port", "sample": sample}
        else:
            out["cpu_baseline"] = None
        out["extra"] = extra
        print(json.dumps(out))
    return 0


if __name__ == "__main__":
    sys.exit(main())
