"""world_size-2 gloo test of the one cross-GPU step: merging the extrema keys with an
integer MAX all-reduce gives every rank the single-process answer, bit for bit, and
row slabs quantised with the merged scale/bias equal the unsharded result."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _f2k(f):
    b = np.asarray(f, np.float32).view(np.int32)
    return b ^ ((b >> 31) & 0x7FFFFFFF)


def _worker(rank, world, port, H, W, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import oracle as O
    from rti_b200 import lights
    from rti_b200.device import key_to_float
    from rti_b200.encoder import allreduce_keys, slab_rows

    L = lights.dome1_lights()
    M = O.pinv(L)
    r0, r1 = slab_rows(H, world, rank)
    st = O.synth_stack(0xABCD, 0, W, H, L, row0=r0, rows=r1 - r0)
    coeffs = O.fit_u8(st, M)
    rgb = O.chroma_avg(st)
    O.lrgb_colour_normalise(rgb, coeffs)
    c = coeffs.reshape(-1, 6)
    keys = torch.from_numpy(np.concatenate([_f2k(-c.min(0)), _f2k(c.max(0))]).astype(np.int32))
    allreduce_keys(keys)
    f = key_to_float(keys.numpy())
    mn, mx = -f[:6], f[6:]
    # quantise the slab with the merged extrema (rti-builder/ptmlib.c:858-880 on the global range)
    scale = ((mx - mn) / np.float32(256.0)).astype(np.float32)
    bias = (np.float32(-256.0) / (mx - mn) * mn).astype(np.int32)
    inv = (np.float32(1.0) / scale).astype(np.float32)
    q = (c * inv).astype(np.float32) + bias.astype(np.float32)
    b = np.where(q > 255, 255, np.where(q < 0, 0, q)).astype(np.uint8)
    ret[rank] = (r0, r1, mn, mx, b.reshape(r1 - r0, W, 6), rgb)
    dist.destroy_process_group()


def test_two_rank_slabs_equal_single(tmp_path):
    H, W, world = 23, 16, 2   # H not divisible by world
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), H, W, ret), nprocs=world, join=True)
    from oracle import oracle as O
    from rti_b200 import lights
    L = lights.dome1_lights()
    full = O.encode_lrgb(O.synth_stack(0xABCD, 0, W, H, L), O.pinv(L))
    got_coef = np.concatenate([ret[r][4] for r in range(world)], axis=0)
    got_rgb = np.concatenate([ret[r][5] for r in range(world)], axis=0)
    for r in range(world):
        np.testing.assert_array_equal(ret[r][2], full["minmax"][:6])
        np.testing.assert_array_equal(ret[r][3], full["minmax"][6:])
    np.testing.assert_array_equal(got_coef, full["coef"])
    np.testing.assert_array_equal(got_rgb, full["rgb"])
