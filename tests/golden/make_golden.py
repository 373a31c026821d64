"""Generate the committed golden fixtures (run in the build container, where
/root/reference exists and oracle/_ref has been built from it):

    python tests/golden/make_golden.py

octave_kat.json   the only numeric example the reference ships
                  (doc_src/draw_light_illustrations.m:1-59): design matrix A, samples b,
                  and X = pinv(A) b evaluated here in float64 with numpy.
dome1_pinv.npz    the 6 x 60 matrix the reference's own ptm_svd (rti-builder/ptmlib.c:630)
                  returns for dome1.lp on this image's OpenBLAS 0.3.15.
lrgb_*.npz / rgb_*.npz
                  outputs of the reference's UNMODIFIED ptm-encoder and ptm-decoder
                  (oracle/_ref/ptm-*-ref, codec replaced by the raw container) on seeded
                  synthetic stacks: scale, bias, all block bytes, relit rasters.
synth_hashes.json sha256 of the synthetic stacks, pinning the generator itself.
"""
import ctypes as C
import hashlib
import json
import os
import re
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from rti_b200 import lights  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"

CASES = {
    # name: (seed, mode, width, height, format)
    "lrgb_64x48": (0x5EED0001, 0, 64, 48, "PTM_FORMAT_LRGB"),
    "lrgb_37x29": (0x5EED0002, 0, 37, 29, "PTM_FORMAT_LRGB"),   # ragged: pixels % 4 != 0
    "rgb_40x24": (0x5EED0003, 1, 40, 24, "PTM_FORMAT_RGB"),
}
DIRECTIONS = [(0.0, 0.0), (0.5, 0.5), (-0.7, 0.3), (0.0, -0.95)]


def octave_kat():
    src = open(os.path.join(REF, "doc_src", "draw_light_illustrations.m")).read()
    a_txt = re.search(r"A = \[(.*?)\];", src, re.S).group(1)
    A = np.array([[float(x) for x in row.split()] for row in a_txt.strip().split(";") if row.strip()])
    b_txt = re.search(r"b = \[(.*?)\]'", src, re.S).group(1)
    b = np.array([float(x) for x in b_txt.split()])
    assert A.shape == (50, 6) and b.shape == (50,)
    U, S, Vt = np.linalg.svd(A, full_matrices=False)
    X = (Vt.T @ np.diag(1.0 / S) @ U.T) @ b
    json.dump(dict(source="doc_src/draw_light_illustrations.m:1-59", A=A.tolist(), b=b.tolist(),
                   X_float64=X.tolist(), singular_values_float64=S.tolist()),
              open(os.path.join(HERE, "octave_kat.json"), "w"), indent=1)
    print("octave X =", X)


def dome1_pinv():
    # the reference's own ptm_svd, called through ctypes.  decoder_t (rti-builder/ptmlib.h:145-151)
    # starts with FILE*, char*, then float u, v, w; ptm_svd reads only u and v.
    class Dec(C.Structure):
        _fields_ = [("fp", C.c_void_p), ("filename", C.c_char_p), ("u", C.c_float), ("v", C.c_float),
                    ("w", C.c_float), ("pad", C.c_char * 256)]
    L = lights.read_skeleton_lp(os.path.join(REF, "dome1.lp"))
    decs = [Dec(None, None, float(u), float(v), float(w)) for u, v, w in L]
    arr = (C.POINTER(Dec) * len(decs))(*[C.pointer(d) for d in decs])
    lib = O.reflib()
    lib.ptm_svd.restype = C.POINTER(C.c_float)
    lib.ptm_svd.argtypes = [C.c_void_p, C.c_int]
    Mp = lib.ptm_svd(arr, len(decs))
    M = np.ctypeslib.as_array(Mp, shape=(6, len(decs))).copy()
    A = np.stack([L[:, 0] ** 2, L[:, 1] ** 2, L[:, 0] * L[:, 1], L[:, 0], L[:, 1], np.ones(len(L), np.float32)], 1)
    sv = O.singular_values(A.astype(np.float32))
    np.savez(os.path.join(HERE, "dome1_pinv.npz"), M=M, lights=L, singular_values=sv,
             blas=np.array(O.lib().orc_blas_corename().decode()))
    print("dome1 M[0,:4] =", M[0, :4], "sv =", sv)


def encodes():
    L = lights.dome1_lights()
    hashes = {}
    for name, (seed, mode, W, H, fmt) in CASES.items():
        stack = O.synth_stack(seed, mode, W, H, L)
        hashes[name] = hashlib.sha256(stack.tobytes()).hexdigest()
        with tempfile.TemporaryDirectory() as d:
            ptm, path = O.ref_encode(stack, L, fmt, d)
            rasters = np.stack([O.ref_decode(path, u, v) for (u, v) in DIRECTIONS])
        out = dict(seed=seed, mode=mode, width=W, height=H, format=fmt, scale_text=ptm["scale"], bias=ptm["bias"],
                   directions=np.array(DIRECTIONS, np.float32), rasters=rasters)
        for i, b in enumerate(ptm["blocks"]):
            out["block%d" % i] = b
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "scale", ptm["scale"], "bias", ptm["bias"])
    # 16-bit generator pin
    s16 = O.synth_stack(0x5EED0004, 1, 24, 16, lights.dome100_lights(), dtype=np.uint16)
    hashes["u16_rgb_24x16_dome100"] = hashlib.sha256(s16.tobytes()).hexdigest()
    json.dump(hashes, open(os.path.join(HERE, "synth_hashes.json"), "w"), indent=1)


if __name__ == "__main__":
    assert O.ref_available(), "build oracle/_ref first (make -C oracle ref)"
    octave_kat()
    dome1_pinv()
    encodes()
