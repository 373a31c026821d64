"""One process, several GPUs: ctypes front end of the C ABI's multi-GPU entries (ptmb_multi_*).

The row-slab sharding lives in the C library (rti_b200/csrc/ptm_multi.cu): one context, one
stream and one issuing host thread per GPU, in-process peer mailboxes over NVLink for the 12
extrema keys.  This module only marshals pointers; torch provides the device buffers.

``bench_multi`` is bench.py's ``--driver c`` arm: the same 24 MP workload as the torchrun arm,
driven from ONE process through ``ptmb_multi_encode_lrgb_dev`` / ``ptmb_multi_encode_host``.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import time

import numpy as np

from . import _lib
from ._lib import NCOEF, U8, U16, ScaleBias
from .device import Context, _ptr


class _BorrowedContext(Context):
    """A context owned by a MultiGpu (never destroyed from Python)."""

    def __init__(self, lib, handle, device):
        self._lib = lib
        self._h = C.c_void_p(handle)
        self.device = device
        self.n_lights = 0

    def close(self):
        self._h = None


class MultiGpu:
    """devices: list of CUDA ordinals (None = all visible GPUs)."""

    def __init__(self, devices=None):
        self._lib = _lib.load()
        h = C.c_void_p()
        if devices is None:
            _lib.check(self._lib.ptmb_multi_create(None, 0, C.byref(h)))
        else:
            arr = (C.c_int * len(devices))(*devices)
            _lib.check(self._lib.ptmb_multi_create(arr, len(devices), C.byref(h)))
        self._h = h
        self.n = self._lib.ptmb_multi_device_count(h)
        self.ctx = []
        for i in range(self.n):
            p = self._lib.ptmb_multi_ctx(h, i)
            self.ctx.append(_BorrowedContext(self._lib, p, self._lib.ptmb_device(p)))
        self.devices = [c.device for c in self.ctx]

    def close(self):
        if getattr(self, "_h", None):
            self._lib.ptmb_multi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def stream(self, i):
        """Raw cudaStream_t of GPU i's context."""
        return self._lib.ptmb_stream(self.ctx[i]._h)

    def slab(self, height, i):
        r0, r1 = C.c_size_t(0), C.c_size_t(0)
        _lib.check(self._lib.ptmb_multi_slab(self._h, height, i, C.byref(r0), C.byref(r1)))
        return r0.value, r1.value

    def host_slab(self, height, i):
        """Rows GPU i takes in encode_host (bandwidth-proportional once calibrated)."""
        r0, r1 = C.c_size_t(0), C.c_size_t(0)
        _lib.check(self._lib.ptmb_multi_host_slab(self._h, height, i, C.byref(r0), C.byref(r1)))
        return r0.value, r1.value

    def set_host_weights(self, weights=None):
        if weights is None:
            _lib.check(self._lib.ptmb_multi_set_host_weights(self._h, None))
        else:
            assert len(weights) == self.n
            arr = (C.c_double * self.n)(*[float(w) for w in weights])
            _lib.check(self._lib.ptmb_multi_set_host_weights(self._h, arr))

    def calibrate_h2d(self, seconds=0.0):
        """Concurrent pinned host -> device probe on every GPU; returns GB/s per GPU and adopts them as weights."""
        out = (C.c_double * self.n)()
        _lib.check(self._lib.ptmb_multi_calibrate_h2d(self._h, float(seconds), out))
        return list(out)

    def set_matrix(self, M):
        M = np.ascontiguousarray(M, dtype=np.float32)
        self._M = M
        _lib.check(self._lib.ptmb_multi_set_matrix(self._h, M.ctypes.data, M.shape[1]))
        for c in self.ctx:
            c.n_lights = M.shape[1]

    def synchronize(self):
        _lib.check(self._lib.ptmb_multi_synchronize(self._h))

    def encode_host(self, stack, M, planes=1, want_coeffs=False, out=None):
        """Whole encode of a HOST stack [n][H][W][3] across the GPUs; same result layout as
        Context.encode_host.  `stack` / `out` may be numpy arrays or pinned torch tensors."""
        n, H, W, _ = stack.shape
        is_np = isinstance(stack, np.ndarray)
        itemsize = stack.dtype.itemsize if is_np else stack.element_size()
        st = U8 if itemsize == 1 else U16
        P = H * W
        if out is not None:
            blocks = out
        elif planes == 1:
            blocks = [np.empty((H, W, NCOEF), np.uint8), np.empty((H, W, 3), np.uint8)]
        else:
            blocks = [np.empty((H, W, NCOEF), np.uint8) for _ in range(3)]
        coeffs = np.empty((planes, P, NCOEF), np.float32) if want_coeffs else None
        sb = ScaleBias()
        M = np.ascontiguousarray(M, dtype=np.float32)
        ptrs = (C.c_void_p * len(blocks))(*[_ptr(b) for b in blocks])
        _lib.check(self._lib.ptmb_multi_encode_host(self._h, _ptr(stack), st, W, H, n, M.ctypes.data, planes, ptrs,
                                                    None if coeffs is None else coeffs.ctypes.data, C.addressof(sb)))
        return dict(blocks=blocks, scale=np.array(sb.scale, dtype=np.float32),
                    bias=np.array(sb.bias, dtype=np.int32), coeffs=coeffs)

    def make_dev_args(self, stacks, coeffs, rgbs, coef_bytes, sbs, sample_type=U8):
        """Pointer arrays for encode_lrgb_dev (built once, reused every step)."""
        n = self.n
        vp = C.c_void_p * n
        sz = C.c_size_t * n
        px = [int(s[0].numel() // 3) for s in stacks]
        return dict(stack=vp(*[s.data_ptr() for s in stacks]),
                    stride=sz(*[p * 3 * stacks[0].element_size() for p in px]), pixels=sz(*px),
                    coeffs=vp(*[t.data_ptr() for t in coeffs]), rgb=vp(*[t.data_ptr() for t in rgbs]),
                    bytes=vp(*[t.data_ptr() for t in coef_bytes]), sb=vp(*[t.data_ptr() for t in sbs]),
                    sample_type=sample_type)

    def encode_lrgb_dev(self, a):
        """K1 -> NVLink key hand-off -> K2 on every GPU's resident slab (asynchronous)."""
        _lib.check(self._lib.ptmb_multi_encode_lrgb_dev(self._h, a["stack"], a["sample_type"], a["stride"],
                                                        a["pixels"], a["coeffs"], a["rgb"], a["bytes"], a["sb"]))


def bench_multi(args, B):
    """bench.py --driver c: ONE process drives all GPUs through the C ABI (B = the bench module)."""
    import torch

    from . import lights as LS
    from .device import light_q14
    from .encoder import PtmEncoder
    from .ptmlib import ptm_svd

    n = args.gpus
    assert torch.cuda.device_count() >= n, "need %d visible GPUs" % n
    W, H = args.width, args.height
    L = LS.dome1_lights()
    M = ptm_svd(L)
    mg = MultiGpu(list(range(n)))
    mg.set_matrix(M)
    stacks, coeffs, rgbs, cbytes, sbs, slabs = [], [], [], [], [], []
    for i in range(n):
        r0, r1 = mg.slab(H, i)
        rows = r1 - r0
        P = rows * W
        dev = torch.device("cuda", mg.devices[i])
        stacks.append(torch.empty((B.N_LIGHTS, rows, W, 3), dtype=torch.uint8, device=dev))
        coeffs.append(torch.empty((P, NCOEF), dtype=torch.float32, device=dev))
        rgbs.append(torch.empty((P, 3), dtype=torch.uint8, device=dev))
        cbytes.append(torch.empty((P, NCOEF), dtype=torch.uint8, device=dev))
        sbs.append(torch.empty(12, dtype=torch.int32, device=dev))
        slabs.append((r0, r1))
        mg.ctx[i].synth(B.SEED, 0, U8, r0 * W, P, light_q14(L), stacks[i], P * 3)
    mg.synchronize()
    for c in mg.ctx:   # the float planes are scratch here too (see bench.py): K2 may drop their L2 lines
        c.set_coeffs_scratch(os.environ.get("PTM_KEEP_COEFFS", "0") != "1")
    a = mg.make_dev_args(stacks, coeffs, rgbs, cbytes, sbs)
    for _ in range(max(args.warmup, 3)):
        mg.encode_lrgb_dev(a)
    mg.synchronize()

    streams = [torch.cuda.ExternalStream(mg.stream(i), device=torch.device("cuda", mg.devices[i])) for i in range(n)]
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    sampler = B.ClockSampler(B.physical_gpu_index(0))
    sampler.start()
    # one untimed encode lines the GPUs up on the device (every K2 waits for all peers' keys) before the start
    # events: the host threads do not release their first launches at the same instant
    mg.encode_lrgb_dev(a)
    for i in range(n):
        with torch.cuda.device(mg.devices[i]):
            ev[i][0].record(streams[i])
    for _ in range(args.steps):
        mg.encode_lrgb_dev(a)
    for i in range(n):
        with torch.cuda.device(mg.devices[i]):
            ev[i][1].record(streams[i])
    mg.synchronize()
    clocks = sampler.stop()
    ms_step = max(e0.elapsed_time(e1) for e0, e1 in ev) / args.steps
    value = W * H / (ms_step * 1e-3) / 1e6

    # byte identity with a single-GPU encode of the whole image (made on GPU 0)
    torch.cuda.set_device(mg.devices[0])
    full = torch.empty((B.N_LIGHTS, H, W, 3), dtype=torch.uint8, device="cuda")
    single = PtmEncoder(W, H, M, planes=1, single=True)
    single.ctx.synth(B.SEED, 0, U8, 0, W * H, light_q14(L), full, W * H * 3)
    single.encode(full)
    torch.cuda.synchronize()
    same = True
    for i, (r0, r1) in enumerate(slabs):
        same = same and torch.equal(single.coef_bytes[0][r0 * W:r1 * W].cpu(), cbytes[i].cpu())
        same = same and torch.equal(single.rgb[r0 * W:r1 * W].cpu(), rgbs[i].cpu())
        same = same and torch.equal(single.sb.cpu(), sbs[i].cpu())
    del full

    # end to end: ONE pinned host stack -> ptmb_multi_encode_host -> ONE set of pinned host blocks
    e2e = None
    if not args.no_e2e:
        host_stack = torch.empty((B.N_LIGHTS, H, W, 3), dtype=torch.uint8).pin_memory()
        for i, (r0, r1) in enumerate(slabs):
            host_stack[:, r0:r1].copy_(stacks[i])
        out = [torch.empty((H, W, 6), dtype=torch.uint8).pin_memory(), torch.empty((H, W, 3), dtype=torch.uint8).pin_memory()]
        if os.environ.get("PTM_E2E_BALANCE", "1") == "1":
            rates = mg.calibrate_h2d()                     # rows follow each GPU's share of the host bandwidth
        else:
            rates = None
            mg.set_host_weights(None)
        host_rows = [mg.host_slab(H, i) for i in range(n)]
        mg.encode_host(host_stack, M, planes=1, out=out)
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            mg.encode_host(host_stack, M, planes=1, out=out)
        e2e_s = (time.perf_counter() - t0) / args.e2e_steps
        ident = (torch.equal(out[0].reshape(-1, 6), single.coef_bytes[0].cpu()) and
                 torch.equal(out[1].reshape(-1, 3), single.rgb.cpu()))
        e2e = {"value": W * H / e2e_s / 1e6, "unit": "Mpixel/s", "ms_per_step": e2e_s * 1e3, "steps": args.e2e_steps,
               "h2d_bytes_per_step": int(host_stack.numel()), "d2h_bytes_per_step": int(out[0].numel() + out[1].numel()),
               "h2d_gbs": host_stack.numel() / e2e_s / 1e9, "identical_to_single_gpu": bool(ident),
               "h2d_share_gbs_per_gpu": rates, "rows_per_gpu": [r1 - r0 for r0, r1 in host_rows],
               "note": "whole-job bytes; one pinned host stack -> ptmb_multi_encode_host -> pinned host blocks"}
    peak, peak_src = B.measured_peak()
    path_gbs = B.ALG_BYTES_PATH * W * H / (ms_step * 1e-3) / 1e9 / n
    print(json.dumps({
        "metric": B.METRIC, "value": value, "unit": "Mpixel/s", "n_gpus": n, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": B.WORKLOAD % (W * H / 1e6, W, H, B.N_LIGHTS), "sharding": "row slabs over %d GPU(s)" % n,
                   "driver": "ONE process, C ABI ptmb_multi_* (one issuing thread per GPU, in-process peer mailboxes)",
                   "l2": "inputs larger than L2, streamed once per step"},
        "roofline": {"bound": "hbm", "kernel": "fit path K1 + K2", "achieved": path_gbs, "peak": peak, "unit": "GB/s",
                     "frac": path_gbs / peak, "algorithmic_bytes_per_px": B.ALG_BYTES_PATH, "peak_source": peak_src,
                     "traffic": None},
        "parity": {"sharded_identical": bool(same)},
        "clocks": clocks, "gpu_launches": args.steps * 2 * n, **({"e2e": e2e} if e2e else {}),
    }))
    mg.close()
