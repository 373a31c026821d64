#!/usr/bin/env python
"""Concurrent pinned host -> device copies on N GPUs of one box (the ceiling of bench.py's e2e leg at N > 1):
one process per GPU, started together, each copying its own pinned buffer; prints per-GPU and aggregate GB/s for
default and write-combined pinned memory, with and without binding the process to the GPU's NUMA-local cores.
    python tools/ubench_h2d_multi.py N [GB per GPU]"""
import ctypes as C
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def worker(gpu, gb, wc, bind, t_start):
    sys.path.insert(0, ROOT)
    import torch
    from rti_b200 import _lib
    if bind:
        sys.path.insert(0, ROOT)
        import bench
        bench.bind_to_gpu_numa_node(gpu)
    torch.cuda.set_device(gpu)
    lib = _lib.load()
    n = int(gb * 1e9)
    p = C.c_void_p()
    _lib.check(lib.ptmb_host_alloc(n, wc, C.byref(p)))
    C.memset(p, 1, n)
    dev = torch.empty(n, dtype=torch.uint8, device="cuda")
    ctx = C.c_void_p()
    _lib.check(lib.ptmb_create(gpu, None, C.byref(ctx)))
    _lib.check(lib.ptmb_upload(ctx, dev.data_ptr(), p, n))
    _lib.check(lib.ptmb_synchronize(ctx))
    while time.time() < t_start:
        time.sleep(0.0005)
    t0 = time.perf_counter()
    reps = 4
    for _ in range(reps):
        _lib.check(lib.ptmb_upload(ctx, dev.data_ptr(), p, n))
    _lib.check(lib.ptmb_synchronize(ctx))
    dt = (time.perf_counter() - t0) / reps
    print("gpu %d wc=%d bind=%d: %.1f GB/s" % (gpu, wc, bind, n / dt / 1e9), flush=True)
    lib.ptmb_host_free(p)


if __name__ == "__main__":
    if len(sys.argv) > 3 and sys.argv[1] == "--worker":
        worker(int(sys.argv[2]), float(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), float(sys.argv[6]))
        sys.exit(0)
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    gb = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    configs = ((0, 0), (1, 0)) if (len(sys.argv) > 3 and sys.argv[3] == "quick") else ((0, 0), (0, 1), (1, 0), (1, 1))
    for wc, bind in configs:
        t_start = time.time() + 20
        procs = [subprocess.Popen([sys.executable, __file__, "--worker", str(g), str(gb), str(wc), str(bind), str(t_start)],
                                  stdout=subprocess.PIPE, text=True) for g in range(N)]
        rates = []
        for pr in procs:
            out = pr.communicate()[0]
            sys.stdout.write(out)
            try:
                rates.append(float(out.strip().split(":")[1].split()[0]))
            except Exception:
                pass
        print("N=%d wc=%d bind=%d aggregate %.1f GB/s" % (N, wc, bind, sum(rates)), flush=True)
