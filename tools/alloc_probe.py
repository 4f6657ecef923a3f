"""How long do device allocations, frees and stream synchronisations take on this box, and how often do they stall?
(developer probe behind DESIGN §4's note on the erratic device-side setup; run under gpurun)"""
import ctypes as C
import time

import numpy as np
import torch

torch.cuda.init()
x = torch.zeros(1 << 20, device="cuda")
rt = C.CDLL("libcudart.so.12")
rt.cudaMalloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
rt.cudaFree.argtypes = [C.c_void_p]


def stats(name, ts):
    ts = np.array(ts) * 1e3
    print(f"{name:28s} n={len(ts):5d}  p50 {np.median(ts):8.3f} ms  p99 {np.percentile(ts, 99):8.3f} ms  max {ts.max():8.3f} ms  "
          f">10ms: {(ts > 10).sum()}  >100ms: {(ts > 100).sum()}  total {ts.sum() / 1e3:.3f} s", flush=True)


for size in (1 << 20, 64 << 20, 512 << 20):
    tm, tf = [], []
    for _ in range(300):
        p = C.c_void_p()
        t0 = time.perf_counter()
        rc = rt.cudaMalloc(C.byref(p), size)
        t1 = time.perf_counter()
        assert rc == 0
        rt.cudaFree(p)
        t2 = time.perf_counter()
        tm.append(t1 - t0)
        tf.append(t2 - t1)
    stats(f"cudaMalloc {size >> 20} MB", tm)
    stats(f"cudaFree   {size >> 20} MB", tf)
ts = []
for _ in range(3000):
    t0 = time.perf_counter()
    x.add_(1.0)
    torch.cuda.synchronize()
    ts.append(time.perf_counter() - t0)
stats("tiny kernel + synchronize", ts)
h = torch.empty(64 << 20, dtype=torch.uint8).pin_memory()
d = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
ts = []
for _ in range(300):
    t0 = time.perf_counter()
    d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    ts.append(time.perf_counter() - t0)
stats("64 MB pinned H2D + sync", ts)
