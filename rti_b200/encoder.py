"""Device-resident PTM encode / relight pipelines (one process per GPU).

The arithmetic is in the CUDA library; this module only owns buffers (torch
tensors), enqueues K1 -> [exchange] -> K2 on torch's current stream and, when a
process group is given, merges the twelve extrema across GPUs with ONE int32
MAX all-reduce over NCCL (the only cross-GPU step of the whole path: every
other step is per pixel).  The keys are order-preserving integer images of the
floats, so the reduction is exact and the sharded result is bit-identical to
the single-GPU one.

Call order mirrors the reference encoder (rti-builder/ptm-encoder.c:307-353,406):
fit -> chroma average -> colour/normalise (all inside K1) -> scale (K2).
"""
from __future__ import annotations

import numpy as np
import torch

from .device import NCOEF, NKEYS, U8, U16, Context, KeyExchange, key_to_float


def slab_rows(height, world_size, rank):
    """Contiguous row slab of rank `rank`: [floor(r*H/G), floor((r+1)*H/G))."""
    r0 = (rank * height) // world_size
    r1 = ((rank + 1) * height) // world_size
    return r0, r1


def allreduce_keys(keys, group=None):
    """Merge extrema keys across ranks (exact: integer MAX)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(keys, op=dist.ReduceOp.MAX, group=group)
    return keys


def make_key_exchange(ctx, group=None):
    """Set up the NVLink peer-memory exchange for the ranks of `group`; returns None (and every
    rank agrees on None) if any rank could not map its peers, so callers fall back to NCCL."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) < 2:
        return None
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    # every rank takes part in both collectives below whatever happened locally, so a failure
    # on one rank can never leave the others waiting
    xch = None
    try:
        xch = KeyExchange(ctx, rank, world)
        mine = xch.handle
    except Exception:
        mine = b""
    handles = [None] * world
    dist.all_gather_object(handles, mine, group=group)
    ok = 1
    if xch is None or any(len(h) != KeyExchange.HANDLE_BYTES for h in handles):
        ok = 0
    else:
        try:
            xch.connect(handles)
        except Exception:
            ok = 0
    flag = torch.tensor([ok], dtype=torch.int32, device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    if int(flag.item()) == 0:
        if xch is not None:
            xch.close()
        return None
    return xch


class PtmEncoder:
    """LRGB (planes=1) or RGB-format (planes=3) encoder for one row slab held in HBM.

    stack: CUDA uint8 tensor [n_lights, rows, width, 3] (stored row order).
    exchange: "p2p" (NVLink peer-memory mailboxes, falling back to NCCL if peers cannot be
    mapped), or "nccl" (int32 MAX all-reduce through torch.distributed)."""

    def __init__(self, width, rows, M, planes=1, device=None, group=None, keep_coeffs=True, exchange="p2p",
                 engine="default", single=False):
        """single=True: this encoder works alone even inside an initialised process group (no exchange)."""
        assert torch.cuda.is_available(), "rti_b200 needs a CUDA device (no CPU fallback)"
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.width, self.rows, self.planes = int(width), int(rows), int(planes)
        self.pixels = self.width * self.rows
        self.group = group
        self.ctx = Context(self.device.index, stream=torch.cuda.current_stream(self.device).cuda_stream)
        self.ctx.set_matrix(M)
        self.ctx.set_fit_engine(engine)
        self.n_lights = self.ctx.n_lights
        P = self.pixels
        dev = self.device
        self.coeffs = torch.empty((self.planes, P, NCOEF), dtype=torch.float32, device=dev)
        self.coef_bytes = torch.empty((self.planes, P, NCOEF), dtype=torch.uint8, device=dev)
        self.rgb = torch.empty((P, 3), dtype=torch.uint8, device=dev) if self.planes == 1 else None
        self.keys = torch.empty(NKEYS, dtype=torch.int32, device=dev)
        self.sb = torch.empty(NKEYS, dtype=torch.int32, device=dev)  # 6 floats + 6 ints, raw
        self.keep_coeffs = bool(keep_coeffs)
        self.ctx.set_coeffs_scratch(not self.keep_coeffs)   # keep_coeffs=False: self.coeffs is scratch after encode()
        self.xchg = make_key_exchange(self.ctx, group) if exchange == "p2p" and not single else None
        import torch.distributed as dist
        multi = not single and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        self.exchange_kind = "none" if not multi else ("p2p" if self.xchg is not None else "nccl")
        # fused LRGB encode: K1 + K2; otherwise keys-init, K1, K2 per plane
        self.launches_per_encode = 2 if (self.planes == 1 and self.exchange_kind in ("none", "p2p")) else 3

    def _use_current_stream(self):
        """Work is enqueued on torch's CURRENT stream of the encoder's device at the time of each call (not
        the one that was current at construction), so it is ordered against the producer of `stack` and
        the consumers of the results."""
        self.ctx.set_stream(torch.cuda.current_stream(self.device).cuda_stream)

    def fit(self, stack):
        """K1 on the slab; leaves float coefficients (+ RGB) and local extrema keys.  `stack` is
        [n_lights, rows, width, 3], contiguous per light; lights may be strided (a row range of a
        larger stack), which is how a slab of a resident image is passed without a copy."""
        assert stack.is_cuda and stack.dim() == 4 and stack[0].is_contiguous()
        self._use_current_stream()
        n = stack.shape[0]
        assert n == self.n_lights and stack[0].numel() == self.pixels * 3
        st = {torch.uint8: U8, torch.uint16: U16}[stack.dtype]
        image_stride = (stack.stride(0) if n > 1 else self.pixels * 3) * stack.element_size()
        self.ctx.keys_init(self.keys)
        if self.planes == 1:
            self.ctx.fit_lrgb(stack, image_stride, self.pixels, self.coeffs, self.rgb, self.keys, st)
        else:
            self.ctx.fit_rgb(stack, image_stride, self.pixels, self.coeffs, self.pixels * NCOEF, self.keys, st)

    def exchange(self):
        self._use_current_stream()
        if self.xchg is not None:
            self.xchg.allreduce_max(self.keys)
        elif self.exchange_kind != "none":
            allreduce_keys(self.keys, self.group)

    def quantise(self):
        """K2: scale/bias from the (global) keys, then bytes for every plane.  The planes are contiguous
        and share one scale/bias (rti-builder/ptmlib.c:867-880), so the RGB formats are one launch over
        planes * pixels coefficient vectors."""
        self._use_current_stream()
        self.ctx.quantise(self.coeffs, self.pixels * self.planes, self.keys, self.coef_bytes, self.sb,
                          consume=not self.keep_coeffs)

    def encode(self, stack):
        """fit -> exchange -> quantise.  LRGB with the peer (or no) exchange goes through the fused
        two-launch entry (ptmb_encode_lrgb_dev); self.keys is not updated on that path."""
        if self.planes == 1 and self.exchange_kind in ("none", "p2p"):
            assert stack.is_cuda and stack.is_contiguous() and stack.shape[0] == self.n_lights
            self._use_current_stream()
            st = {torch.uint8: U8, torch.uint16: U16}[stack.dtype]
            self.ctx.encode_lrgb_dev(stack, self.pixels * 3 * stack.element_size(), self.pixels, self.coeffs,
                                     self.rgb, self.coef_bytes, self.sb, xchg=self.xchg, sample_type=st)
            return
        if self.planes == 3 and self.exchange_kind in ("none", "p2p") and stack.is_contiguous():
            # RGB formats: tensor-core K1 with the key hand-off + one K2 over the three planes (two launches)
            self._use_current_stream()
            st = {torch.uint8: U8, torch.uint16: U16}[stack.dtype]
            self.ctx.encode_rgb_dev(stack, self.pixels * 3 * stack.element_size(), self.pixels, self.coeffs,
                                    self.coef_bytes, self.sb, xchg=self.xchg, sample_type=st)
            return
        self.fit(stack)
        self.exchange()
        self.quantise()

    # ---- results (synchronising)
    def scale_bias(self):
        raw = self.sb.cpu().numpy()
        return raw[:NCOEF].view(np.float32).copy(), raw[NCOEF:].copy()

    def minmax(self):
        f = key_to_float(self.keys.cpu().numpy())
        return -f[:NCOEF], f[NCOEF:]


class PtmRelighter:
    """Batched relighting of a PTM held in HBM (K3)."""

    def __init__(self, width, height, device=None):
        assert torch.cuda.is_available(), "rti_b200 needs a CUDA device (no CPU fallback)"
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.width, self.height = int(width), int(height)
        self.ctx = Context(self.device.index, stream=torch.cuda.current_stream(self.device).cuda_stream)

    def _use_current_stream(self):
        self.ctx.set_stream(torch.cuda.current_stream(self.device).cuda_stream)

    def relight_lrgb(self, coef, rgb, scale, bias, uv, out=None):
        self._use_current_stream()
        uv = np.asarray(uv, dtype=np.float32).reshape(-1, 2)
        if out is None:
            out = torch.empty((uv.shape[0], self.height, self.width, 3), dtype=torch.uint8, device=self.device)
        self.ctx.relight_lrgb(coef, rgb, self.width, self.height, scale, bias, uv, out)
        return out

    def relight_rgb(self, r, g, b, scale, bias, uv, out=None):
        self._use_current_stream()
        uv = np.asarray(uv, dtype=np.float32).reshape(-1, 2)
        if out is None:
            out = torch.empty((uv.shape[0], self.height, self.width, 3), dtype=torch.uint8, device=self.device)
        self.ctx.relight_rgb(r, g, b, self.width, self.height, scale, bias, uv, out)
        return out
