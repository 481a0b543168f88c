"""Generates tests/golden/ref_collinear_v1.npz by running the UNMODIFIED reference find_collinear_segments_GPU /
K_collinearity (cudawrapper.cu:370-429, 689-705; oracle/_ref built with -fmad=false) on a B200:

    gpurun -- python tests/golden/make_golden_collinear.py gpurun_out/ref_collinear_v1.npz    # then copy into tests/golden/

Stored per case: the segments and the (row, col) indices of the ones of the N x N char matrix.  The CPU suite checks
oracle/l3d_oracle.cc (orc_collinear_f32) against it; the GPU suite checks the product against the same reference live.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from line3dpp_b200 import synth  # noqa: E402
from oracle import pyoracle as po  # noqa: E402


def edge_case_segments():
    """degenerate and adversarial inputs: zero-length segments, exact duplicates, exactly collinear chains with and
    without gaps, touching endpoints, reversed direction, far-away coordinates"""
    s = [
        (100, 100, 200, 100), (210, 100, 300, 100), (300, 100, 400, 100), (250, 100, 350, 100),     # chain, touching, overlapping
        (400, 100.5, 500, 101.0), (600, 101.5, 500.5, 101.0), (100, 100, 200, 100),                 # nearly collinear, reversed, duplicate
        (50, 50, 50, 50), (50, 50, 50, 50), (60, 60, 60, 60.0001),                                   # zero / tiny length
        (0, 0, 3071, 2303), (1, 1, 10, 7.75), (1000, 750.2, 2000, 1500.1), (2500, 1875, 3000, 2250),  # long diagonal and pieces on it
        (100, 2000, 100, 2100), (100, 2150, 100, 2250), (101.5, 1800, 101.9, 1950),                  # vertical
        (1e6, 1e6, 1e6 + 50, 1e6), (1e6 + 80, 1e6, 1e6 + 150, 1e6), (-500, -500, -400, -500),       # far away
    ]
    return np.array(s, np.float32)


def main(out_path):
    ref = po.ref_lib("nofma")
    assert ref is not None and ref.ref_device_count() > 0, "needs oracle/_ref and a GPU"
    sc = synth.make_scene(4, 220, 91, "ring1", collinear=True)
    g = {"scene_args": np.array([4, 220, 91])}
    cases = {"v0": sc.segs[0], "v2": sc.segs[2], "edge": edge_case_segments()}
    for name, segs in cases.items():
        g[f"{name}_segs"] = segs
        for t in (0.5, 2.0, 6.0):
            Cm, _ = po.collinear(ref.ref_collinear, segs, t)
            assert np.array_equal(Cm, Cm.T) and set(np.unique(Cm)) <= {0, 1}
            g[f"{name}_t{t}"] = np.argwhere(Cm == 1).astype(np.int32)
    np.savez_compressed(out_path, **g)
    print("wrote", out_path, {k: v.shape for k, v in g.items()})


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/ref_collinear_v1.npz")
