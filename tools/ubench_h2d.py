#!/usr/bin/env python
"""Host -> device copy ceilings on this box (one GPU per process; run one process per GPU for the concurrent case):
pinned cudaMemcpyAsync, pageable cudaMemcpy, cudaHostRegister cost, threaded memcpy into a pinned staging ring.
    python tools/ubench_h2d.py [GB] [gpu]"""
import ctypes
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

gb = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
gpu = int(sys.argv[2]) if len(sys.argv) > 2 else 0
torch.cuda.set_device(gpu)
n = int(gb * 1e9)
dev = torch.empty(n, dtype=torch.uint8, device="cuda")
rt = torch.cuda.cudart()


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


pinned = torch.empty(n, dtype=torch.uint8).pin_memory()
pinned.fill_(1)
t = timed(lambda: dev.copy_(pinned, non_blocking=True))
print("pinned H2D            %6.1f GB/s" % (n / t / 1e9))
t = timed(lambda: pinned.copy_(dev, non_blocking=True))
print("pinned D2H            %6.1f GB/s" % (n / t / 1e9))
pageable = np.ones(n, np.uint8)
tp = torch.from_numpy(pageable)
t = timed(lambda: dev.copy_(tp))
print("pageable H2D          %6.1f GB/s" % (n / t / 1e9))
back = torch.from_numpy(np.empty(n, np.uint8))
back.fill_(0)
t = timed(lambda: back.copy_(dev))
print("pageable D2H          %6.1f GB/s" % (n / t / 1e9))
t0 = time.perf_counter()
rc = rt.cudaHostRegister(pageable.ctypes.data, n, 0)
t1 = time.perf_counter()
print("cudaHostRegister      %6.1f GB/s (%.3f s for %.1f GB, rc %s)" % (n / (t1 - t0) / 1e9, t1 - t0, gb, rc))
t = timed(lambda: dev.copy_(tp, non_blocking=True))
print("registered H2D        %6.1f GB/s" % (n / t / 1e9))
t0 = time.perf_counter()
rt.cudaHostUnregister(pageable.ctypes.data)
print("cudaHostUnregister    %.3f s" % (time.perf_counter() - t0))
# threaded memcpy pageable -> pinned (what a staging ring would do), no GPU involved
chunk = 16 << 20
for threads in (1, 4, 8, 16, 32):
    if threads > (os.cpu_count() or 1):
        break
    pool = ThreadPoolExecutor(threads)
    src = pageable
    dst = pinned.numpy()
    parts = [(i, min(i + chunk, n)) for i in range(0, n, chunk)]

    def copy(p):
        ctypes.memmove(dst.ctypes.data + p[0], src.ctypes.data + p[0], p[1] - p[0])

    list(pool.map(copy, parts))
    t0 = time.perf_counter()
    list(pool.map(copy, parts))
    t = time.perf_counter() - t0
    print("memcpy -> pinned, %2d threads %6.1f GB/s" % (threads, n / t / 1e9))
    pool.shutdown()
print("cpus", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
