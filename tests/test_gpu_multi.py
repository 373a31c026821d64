"""Multi-GPU test (needs >= 2 GPUs; skipped otherwise): two ranks, row slabs of one image, the
peer-memory key exchange and the NCCL all-reduce give the same keys, and the sharded blocks are
byte-identical to the single-GPU encode."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, H, W, ret):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from oracle import oracle as O
    from rti_b200 import lights
    from rti_b200.encoder import PtmEncoder, slab_rows
    L = lights.dome1_lights()
    M = O.pinv(L)
    r0, r1 = slab_rows(H, world, rank)
    slab = torch.from_numpy(O.synth_stack(0xD157, 0, W, H, L, row0=r0, rows=r1 - r0)).cuda()
    out = {}
    for kind in ("p2p", "nccl"):
        enc = PtmEncoder(W, r1 - r0, M, exchange=kind)
        for _ in range(3):                       # several epochs through the mailboxes
            enc.encode(slab)
        torch.cuda.synchronize()
        status = enc.xchg.status() if enc.xchg is not None else 0
        # the fused path leaves no keys; scale / bias (derived from the merged keys) stand in
        out[kind] = (enc.exchange_kind, status, enc.sb.cpu().numpy(), enc.coef_bytes[0].cpu().numpy(),
                     enc.rgb.cpu().numpy())
        # stepwise calls through the same exchange still work after fused ones
        enc.fit(slab)
        enc.exchange()
        enc.quantise()
        torch.cuda.synchronize()
        assert np.array_equal(enc.sb.cpu().numpy(), out[kind][2])
    ret[rank] = (r0, r1, out)
    dist.destroy_process_group()


def test_two_gpu_slabs_equal_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    H, W, world = 301, 128, 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), H, W, ret), nprocs=world, join=True)
    from oracle import oracle as O
    from rti_b200 import lights
    from rti_b200.encoder import PtmEncoder
    L = lights.dome1_lights()
    M = O.pinv(L)
    torch.cuda.set_device(0)
    full = PtmEncoder(W, H, M)
    full.encode(torch.from_numpy(O.synth_stack(0xD157, 0, W, H, L)).cuda())
    torch.cuda.synchronize()
    kinds = {ret[r][2]["p2p"][0] for r in range(world)}
    print("exchange used:", kinds)
    for kind in ("p2p", "nccl"):
        for r in range(world):
            assert ret[r][2][kind][1] == 0
            np.testing.assert_array_equal(ret[r][2][kind][2], full.sb.cpu().numpy())
        coef = np.concatenate([ret[r][2][kind][3] for r in range(world)])
        rgb = np.concatenate([ret[r][2][kind][4] for r in range(world)])
        np.testing.assert_array_equal(coef, full.coef_bytes[0].cpu().numpy())
        np.testing.assert_array_equal(rgb, full.rgb.cpu().numpy())


# ---------------------------------------------------------------- one process, C ABI (ptmb_multi_*)

def _single_gpu_reference(oracle_mod, W, H, planes, seed, mode):
    from rti_b200 import lights
    from rti_b200.device import Context
    L = lights.dome1_lights()
    M = oracle_mod.pinv(L)
    stack = oracle_mod.synth_stack(seed, mode, W, H, L)
    torch.cuda.set_device(0)
    c = Context(0, stream=torch.cuda.current_stream().cuda_stream)
    want = c.encode_host(stack, M, planes=planes, want_coeffs=True)
    c.close()
    return L, M, stack, want


@pytest.mark.parametrize("planes,mode", [(1, 0), (3, 1)])
def test_multi_encode_host_equals_single_gpu(oracle, planes, mode):
    """ptmb_multi_encode_host over every visible GPU (and over GPU 0 alone, the degenerate case that also runs on
    a one-GPU box): blocks, float coefficients and scale / bias byte-identical to the single-GPU call.  Height
    301 does not divide by the GPU count; width 128 gives slabs that are not tile multiples."""
    from rti_b200.multi import MultiGpu
    W, H = 128, 301
    L, M, stack, want = _single_gpu_reference(oracle, W, H, planes, 0xD157, mode)
    for devices in ([0], list(range(torch.cuda.device_count()))):
        if len(devices) == 1 and devices != [0]:
            continue
        mg = MultiGpu(devices)
        for _ in range(2):   # twice: mailbox epochs and scratch reuse
            got = mg.encode_host(stack, M, planes=planes, want_coeffs=True)
        for a, b in zip(got["blocks"], want["blocks"]):
            np.testing.assert_array_equal(a, b)
        np.testing.assert_array_equal(got["coeffs"], want["coeffs"])
        np.testing.assert_array_equal(got["scale"], want["scale"])
        np.testing.assert_array_equal(got["bias"], want["bias"])
        print("ptmb_multi_encode_host on GPUs %s: identical to the single-GPU result" % devices)
        mg.close()


def test_multi_encode_host_unequal_slabs(oracle):
    """Bandwidth-proportional slabs (ptmb_multi_calibrate_h2d / ptmb_multi_set_host_weights): any split of the rows
    gives the single-GPU bytes, including a split that leaves a GPU without rows."""
    from rti_b200.multi import MultiGpu
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs 2 GPUs")
    W, H = 128, 301
    L, M, stack, want = _single_gpu_reference(oracle, W, H, 1, 0xD159, 0)
    mg = MultiGpu(None)
    rates = mg.calibrate_h2d(0.005)
    assert len(rates) == n and all(r > 0.5 for r in rates), rates      # GB/s
    print("concurrent pinned H2D, GB/s per GPU:", ["%.1f" % r for r in rates])
    for weights in (None, [1.0 + 2.0 * (i % 2) for i in range(n)], [1e-9] + [1.0] * (n - 1), "calibrated"):
        if weights == "calibrated":
            mg.calibrate_h2d(0.005)
        else:
            mg.set_host_weights(weights)
        cuts = [mg.host_slab(H, i) for i in range(n)]
        assert cuts[0][0] == 0 and cuts[-1][1] == H and all(cuts[i][1] == cuts[i + 1][0] for i in range(n - 1)), cuts
        got = mg.encode_host(stack, M, planes=1, want_coeffs=True)
        for a, b in zip(got["blocks"], want["blocks"]):
            np.testing.assert_array_equal(a, b)
        np.testing.assert_array_equal(got["coeffs"], want["coeffs"])
        np.testing.assert_array_equal(got["bias"], want["bias"])
    mg.close()


def test_multi_more_gpus_than_rows(oracle):
    """Slabs may be empty (2 rows over up to 8 GPUs): empty GPUs still take part in the key exchange."""
    from rti_b200.multi import MultiGpu
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    W, H = 64, 2
    L, M, stack, want = _single_gpu_reference(oracle, W, H, 1, 0xD158, 0)
    mg = MultiGpu(None)
    got = mg.encode_host(stack, M, planes=1)
    for a, b in zip(got["blocks"], want["blocks"]):
        np.testing.assert_array_equal(a, b)
    np.testing.assert_array_equal(got["bias"], want["bias"])
    mg.close()


def test_multi_encode_lrgb_dev_equals_single_gpu(oracle):
    """Device-resident slabs through ptmb_multi_encode_lrgb_dev (K1 -> in-process NVLink mailboxes -> K2)."""
    from rti_b200 import lights
    from rti_b200.device import NCOEF
    from rti_b200.encoder import PtmEncoder
    from rti_b200.multi import MultiGpu
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    W, H = 1024, 96 * torch.cuda.device_count() + 5
    L = lights.dome1_lights()
    M = oracle.pinv(L)
    stack = oracle.synth_stack(0xD159, 0, W, H, L)
    torch.cuda.set_device(0)
    full = PtmEncoder(W, H, M)
    full.encode(torch.from_numpy(stack).cuda())
    torch.cuda.synchronize()
    mg = MultiGpu(None)
    mg.set_matrix(M)
    stacks, coeffs, rgbs, cbytes, sbs, slabs = [], [], [], [], [], []
    for i in range(mg.n):
        r0, r1 = mg.slab(H, i)
        dev = torch.device("cuda", mg.devices[i])
        P = (r1 - r0) * W
        stacks.append(torch.from_numpy(np.ascontiguousarray(stack[:, r0:r1])).to(dev))
        coeffs.append(torch.empty((P, NCOEF), dtype=torch.float32, device=dev))
        rgbs.append(torch.empty((P, 3), dtype=torch.uint8, device=dev))
        cbytes.append(torch.empty((P, NCOEF), dtype=torch.uint8, device=dev))
        sbs.append(torch.empty(12, dtype=torch.int32, device=dev))
        slabs.append((r0, r1))
    for i in range(mg.n):
        torch.cuda.synchronize(mg.devices[i])
    a = mg.make_dev_args(stacks, coeffs, rgbs, cbytes, sbs)
    for _ in range(5):
        mg.encode_lrgb_dev(a)
    mg.synchronize()
    for i, (r0, r1) in enumerate(slabs):
        assert torch.equal(cbytes[i].cpu(), full.coef_bytes[0][r0 * W:r1 * W].cpu())
        assert torch.equal(rgbs[i].cpu(), full.rgb[r0 * W:r1 * W].cpu())
        assert torch.equal(sbs[i].cpu(), full.sb.cpu())
    mg.close()


def test_ptm_encode_stack_uses_every_listed_gpu(oracle, tmp_path):
    """The ptmlib.h drop-in with PTM_B200_DEVICES=0,1,...: ptm_encode_stack shards the rows; the PTM file written
    by the command line is byte-identical to the single-GPU one."""
    import subprocess
    from conftest import ROOT
    from rti_b200 import lights
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from test_gpu_cli import write_inputs
    L = lights.dome1_lights()
    W, H = 96, 50
    stack = oracle.synth_stack(0xC12, 0, W, H, L)
    lp = write_inputs(oracle, str(tmp_path), stack, L)
    enc = os.path.join(ROOT, "rti_b200", "bin", "ptm-encoder")
    outs = []
    for devs in ("0", ",".join(str(i) for i in range(torch.cuda.device_count()))):
        out = os.path.join(str(tmp_path), "out_%d.ptm" % len(outs))
        env = dict(os.environ, PTM_IMAGE_CODEC="raw", OPENBLAS_NUM_THREADS="1")
        env.pop("CUDA_VISIBLE_DEVICES", None)
        env["PTM_B200_DEVICES" if "," in devs else "PTM_B200_DEVICE"] = devs
        r = subprocess.run([enc, "-f", "PTM_FORMAT_LRGB", "-o", out, lp], capture_output=True, env=env)
        assert r.returncode == 0, r.stderr.decode()
        outs.append(open(out, "rb").read())
    assert outs[0] == outs[1]
