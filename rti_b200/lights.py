"""Light-position sets and the ``.lp`` grammar.

The reference's light set for every headline configuration is ``dome1.lp``
(/root/reference/dome1.lp: 60 lines ``u v w``), which its generator script
derives from the ring geometry of the prototype dome
(/root/reference/scripts/lp-builder.py:18-19,38-52): rings of 4+8+16+16+16
LEDs, each ring starting at 45 degrees and running clockwise, positions
``(d sin(phi), d cos(phi), sqrt(1 - d^2))`` printed with ``%f``.

``ring_lights`` restates that recipe so the same positions (and larger sets
such as the 100-light configuration) can be produced without the reference.
``read_lp`` follows the line grammar of the reference's encoder
(/root/reference/rti-builder/ptm-encoder.c:152-157): a line counts when it has
at least ``name u v`` -- anything else (e.g. a leading count line) is skipped.
"""
from __future__ import annotations

import math
import os

import numpy as np

# (number of LEDs, ring diameter in mm); inner dome diameter in mm
DOME1_RINGS = ((4, 85.0), (8, 218.0), (16, 330.0), (16, 416.0), (16, 457.0))
DOME1_DIAMETER = 490.0 - 2 * 13.0

# A fixed 100-light arrangement for the 50 MP / 16-bit configuration (BASELINE.json
# configs[3]); same recipe, more LEDs per ring.
DOME100_RINGS = ((4, 85.0), (8, 160.0), (16, 218.0), (24, 330.0), (24, 416.0), (24, 457.0))

_DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


def ring_lights(rings=DOME1_RINGS, dome_diameter=DOME1_DIAMETER, angle_deg=0.0, decimals=6):
    """Light directions from ring geometry -> float32 array [n, 3] (u, v, w).

    ``decimals=6`` reproduces the ``%f`` rounding of the generated file."""
    out = []
    phi0 = 2.0 * math.pi * 45.0 / 360.0 + 2.0 * math.pi * angle_deg / 360.0
    for count, diameter in rings:
        d = diameter / dome_diameter
        w = math.sqrt(1.0 - d * d)
        phi = phi0
        for _ in range(count):
            row = (d * math.sin(phi), d * math.cos(phi), w)
            if decimals is not None:
                row = tuple(float("%.*f" % (decimals, x)) for x in row)
            out.append(row)
            phi -= 2.0 * math.pi / count
    return np.asarray(out, dtype=np.float32)


def dome1_lights():
    """The 60 dome1.lp directions, float32 [60, 3], in file order."""
    return ring_lights(DOME1_RINGS)


def dome100_lights():
    """The 100-light set used for the 16-bit configuration."""
    return ring_lights(DOME100_RINGS)


def dome1_lp_path():
    return os.path.join(_DATA_DIR, "dome1.lp")


def read_lp(path):
    """Parse an ``.lp`` file -> (names, float32 [n, 3]).

    Accepts ``name u v [w]`` lines; lines with fewer than three fields are
    ignored, like the encoder's ``sscanf("%s %f %f %f") >= 3`` test."""
    names, rows = [], []
    with open(path, "r") as fp:
        for line in fp:
            parts = line.split()
            if len(parts) < 3:
                continue
            try:
                u, v = float(parts[1]), float(parts[2])
            except ValueError:
                continue
            w = 0.0
            if len(parts) > 3:
                try:
                    w = float(parts[3])
                except ValueError:
                    w = 0.0
            names.append(parts[0])
            rows.append((u, v, w))
    return names, np.asarray(rows, dtype=np.float32).reshape(-1, 3)


def read_skeleton_lp(path):
    """Parse a skeleton file of bare ``u v w`` lines (dome1.lp itself)."""
    rows = []
    with open(path, "r") as fp:
        for line in fp:
            parts = line.split()
            if len(parts) == 3:
                rows.append(tuple(float(x) for x in parts))
    return np.asarray(rows, dtype=np.float32).reshape(-1, 3)


def write_lp(path, names, lights):
    """Write ``count`` + ``name u v w`` lines (the layout image-filer.py emits,
    /root/reference/scripts/image-filer.py:109-115)."""
    with open(path, "w") as fp:
        fp.write("%d\n" % len(names))
        for name, (u, v, w) in zip(names, lights):
            fp.write("%s %f %f %f\n" % (name, u, v, w))
