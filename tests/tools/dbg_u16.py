import sys, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from oracle import oracle as O
from rti_b200.device import Context, U16
def ring(n):
    th = np.arange(n) * 2.399963
    r = 0.15 + 0.8 * (np.arange(n) + 0.5) / n
    u, v = r * np.cos(th), r * np.sin(th)
    return np.stack([u, v, np.sqrt(1 - u * u - v * v)], 1).astype(np.float32)
ctx = Context(0); ctx.set_fit_engine("mma")
for (W, H, n) in [(48, 5, 224), (64, 4, 224), (48, 5, 192), (48, 5, 160), (48,5,128)]:
    L = ring(n); M = O.pinv(L); P = W * H
    s16 = O.synth_stack(0x6000 + n + W, 1, W, H, L, dtype=np.uint16)
    want = O.encode_rgb_u16(s16, M)["coeffs"].reshape(3, P, 6)
    ctx.set_matrix(M)
    coeffs = torch.zeros((3, P, 6), dtype=torch.float32, device="cuda")
    keys = torch.empty(12, dtype=torch.int32, device="cuda"); ctx.keys_init(keys)
    b = ctx.mma_launches()
    ctx.fit_rgb(torch.from_numpy(s16.view(np.int16)).cuda().view(torch.uint16), P * 6, P, coeffs, P * 6, keys, U16)
    torch.cuda.synchronize()
    got = coeffs.cpu().numpy()
    bound = np.einsum("kn,npc->cpk", np.abs(M).astype(np.float64), s16.reshape(n, P, 3).astype(np.float64))
    err = np.abs(got - want) / np.maximum(bound, 1e-30)
    bad = np.argwhere(err > 1e-5)
    print(W, H, n, "mma ran", ctx.mma_launches() - b, "max err", err.max(), "bad count", len(bad), "of", err.size)
    if len(bad):
        print("  bad planes", np.unique(bad[:, 0]), "coefs", np.unique(bad[:, 2]), "px range", bad[:, 1].min(), bad[:, 1].max())
        print("  first bad", bad[:8].tolist())
        c, p, k = bad[0]
        print("  got", got[c, p], "want", want[c, p])
