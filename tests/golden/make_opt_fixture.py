"""Builds tests/golden/line3dpp_ref_opt_pairs_v1.npz from the reference's own result fixtures (run in the build container,
where /root/reference exists):

    testdata/Line3D++_ref/Line3D++__...__vis_3.txt              result WITHOUT Ceres bundling
    testdata/Line3D++_ref/Line3D++__...__OPTIMIZED__vis_3.txt   result of the same configuration WITH it (optimization.cc)

Both list, per final 3D line, its collinear 3D segments and its 2D residuals (camID segID x1 y1 x2 y2, README.md:272-277).
A cluster whose residual set is identical in both files went through LineOptimizer::optimize unchanged in membership, so
(line before, residuals) -> (line after) is a known-answer vector of the reference's OWN Ceres run: the input of the
optimiser is the infinite line through the un-optimised segments, the expected output the line through the optimised
ones.  Stored per pair: a point + unit direction of both lines, and the residuals (cam, seg, x1, y1, x2, y2).
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_nvm_inputs import read_fixture  # noqa: E402

REF = "/root/reference/testdata/Line3D++_ref"
BASE = "Line3D++__W_FULL__N_10__sigmaP_2.5__sigmaA_10__epiOverlap_0.25__kNN_10__"
OUT = os.path.dirname(os.path.abspath(__file__))


def lines_of(segs, seg_line, res):
    out = {}
    for ln in sorted(set(seg_line.tolist())):
        S = segs[seg_line == ln]
        pts = np.concatenate([S[:, :3], S[:, 3:]])
        c = pts.mean(0)
        _, _, vt = np.linalg.svd(pts - c)
        d = vt[0]
        t = (pts - c) @ d
        r = res[res[:, 0] == ln][:, 1:]
        key = tuple(sorted((int(a), int(b)) for a, b in r[:, :2]))
        out[key] = (c + d * t.min(), c + d * t.max(), r, float(np.abs((pts - c) - np.outer(t, d)).max()))
    return out


def main():
    a = lines_of(*read_fixture(os.path.join(REF, BASE + "vis_3.txt")))
    b = lines_of(*read_fixture(os.path.join(REF, BASE + "OPTIMIZED__vis_3.txt")))
    keys = sorted(set(a) & set(b))
    before = np.array([np.concatenate(a[k][:2]) for k in keys])
    after = np.array([np.concatenate(b[k][:2]) for k in keys])
    ptr = np.concatenate([[0], np.cumsum([len(a[k][2]) for k in keys])]).astype(np.int64)
    res = np.concatenate([a[k][2] for k in keys])
    print(len(a), "lines,", len(b), "optimised lines,", len(keys), "with identical residual sets;", len(res), "residuals; max off-line",
          max(a[k][3] for k in keys), max(b[k][3] for k in keys))
    np.savez_compressed(os.path.join(OUT, "line3dpp_ref_opt_pairs_v1.npz"), before=before, after=after, res_ptr=ptr, residuals=res)


if __name__ == "__main__":
    main()
