"""Scratch micro-benchmark of the GEMV kernel (not the judged bench; see bench.py)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from merizo_search_b200 import native, synth

def run(n, nq, k, iters=50):
    dev = torch.device("cuda:0")
    h = native.Database(n)
    for b, r0 in enumerate(range(0, n, 1 << 20)):
        r = min(1 << 20, n - r0)
        x = synth.device_block(b, r, dev)
        h.upload_device(r0, r, x.data_ptr())
        del x
    h.finalize()
    q = torch.nn.functional.normalize(torch.randn(nq, 128, device=dev))
    sc = torch.empty(nq, k, device=dev); ids = torch.empty(nq, k, dtype=torch.int64, device=dev)
    stream = torch.cuda.Stream()
    torch.cuda.synchronize()
    st = stream.cuda_stream
    for _ in range(5):
        h.search_device(q.data_ptr(), nq, k, sc.data_ptr(), ids.data_ptr(), mode=native.MODE_GEMV, stream=st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record()
        for _ in range(iters):
            h.search_device(q.data_ptr(), nq, k, sc.data_ptr(), ids.data_ptr(), mode=native.MODE_GEMV, stream=st)
        e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    gbs = n * 512 / ms / 1e6
    # e2e host call
    qh = q.cpu().numpy()
    t0 = time.perf_counter()
    for _ in range(iters):
        h.search(qh, k, mode=native.MODE_GEMV)
    e2e = (time.perf_counter() - t0) / iters * 1e3
    print(f"n={n} nq={nq} k={k}: {ms*1e3:.1f} us/launch  {gbs:.0f} GB/s ({gbs/6452.8*100:.1f}% of measured HBM)  e2e host call {e2e*1e3:.1f} us", flush=True)
    h.close()

if __name__ == "__main__" and len(sys.argv) == 1:
    for n, nq, k in [(500000, 1, 10), (500000, 1, 100), (500000, 2, 10), (500000, 4, 10), (500000, 8, 10), (500000, 8, 100), (4000000, 1, 10), (4000000, 8, 10), (40000000, 1, 10)]:
        run(n, nq, k)


def upload_speed(n=4_000_000):
    """Host -> device load rate of fcs_db_upload (pageable numpy rows through the pinned staging ring)."""
    x = np.random.default_rng(0).standard_normal((n, 128), dtype=np.float32)
    h = native.Database(n)
    t0 = time.perf_counter()
    for r0 in range(0, n, 262144):
        h.upload(r0, x[r0:r0 + 262144])
    h.finalize()
    dt = time.perf_counter() - t0
    print(f"upload+finalize {n} rows ({n*512/1e9:.2f} GB): {dt:.3f} s = {n*512/1e9/dt:.1f} GB/s", flush=True)
    h.close()


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "upload":
    upload_speed()
