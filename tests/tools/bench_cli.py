#!/usr/bin/env python
"""Wall-clock of the command lines on BASELINE.json configs[0]: 1024 x 1024 x 60 lights (dome1.lp),
LRGB, pre-decoded inputs in the raw container -- the reference's unmodified ptm-encoder / ptm-decoder
(oracle/_ref, all host cores) against rti_b200/bin (one B200, process start and CUDA context
creation included).  Also checks that both encoders' files agree."""
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from rti_b200 import lights  # noqa: E402


def timed(cmd, **kw):
    t0 = time.perf_counter()
    r = subprocess.run(cmd, capture_output=True, env=dict(os.environ, OPENBLAS_NUM_THREADS="1", PTM_IMAGE_CODEC="raw"), **kw)
    assert r.returncode == 0, r.stderr.decode()[-400:]
    return time.perf_counter() - t0, r


def main():
    W = H = 1024
    L = lights.dome1_lights()
    stack = O.synth_stack(0x5EED0001, 0, W, H, L)
    with tempfile.TemporaryDirectory() as d:
        for i in range(60):
            O.write_rawj(os.path.join(d, "img%03d.rawj" % i), stack[i, ::-1])
        lp = os.path.join(d, "s.lp")
        with open(lp, "w") as fp:
            fp.write("60\n")
            for i, (u, v, w) in enumerate(L):
                fp.write("img%03d.rawj %.9g %.9g %.9g\n" % (i, u, v, w))
        out = {}
        for name, enc, dec in (("reference", os.path.join(ROOT, "oracle/_ref/ptm-encoder-ref"), os.path.join(ROOT, "oracle/_ref/ptm-decoder-ref")),
                               ("b200", os.path.join(ROOT, "rti_b200/bin/ptm-encoder"), os.path.join(ROOT, "rti_b200/bin/ptm-decoder"))):
            ptm = os.path.join(d, name + ".ptm")
            te = min(timed([enc, "-f", "PTM_FORMAT_LRGB", "-o", ptm, lp])[0] for _ in range(3))
            td = min(timed([dec, ptm, "0.3", "0.2"])[0] for _ in range(3))
            out[name] = {"encode_s": te, "decode_s": td, "encode_Mpixel_per_s": W * H / te / 1e6}
        a = np.frombuffer(open(os.path.join(d, "reference.ptm"), "rb").read(), np.uint8)
        b = np.frombuffer(open(os.path.join(d, "b200.ptm"), "rb").read(), np.uint8)
        out["files"] = {"same_size": bool(a.size == b.size),
                        "differing_bytes": int((a != b).sum()) if a.size == b.size else None, "bytes": int(a.size)}
        out["host_cores"] = os.cpu_count()
        print(json.dumps(out))


if __name__ == "__main__":
    main()
