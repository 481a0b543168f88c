"""Helpers for the testdata/vsfm_result.nvm configuration (BASELINE.json configs[0]/[1]): load the committed inputs and
compare a reconstruction with the reference's own result fixture statistically (segment ids cannot line up)."""
import os
import numpy as np

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_inputs():
    z = np.load(os.path.join(G, "nvm_inputs_v1.npz"))
    V = len(z["K"])
    return dict(V=V, K=z["K"], R=z["R"], t=z["t"], median_depth=z["median_depth"], wh=z["wh"],
                segs=[z[f"segs_{i}"] for i in range(V)], wps=[z[f"wps_{i}"] for i in range(V)])


def load_fixture():
    z = np.load(os.path.join(G, "line3dpp_ref_fixture_v1.npz"))
    return z["segs3d"], z["seg_line"], z["residuals"]


def add_all(pipe, inp):
    for i in range(inp["V"]):
        w, h = inp["wh"][i]
        rc = pipe(i, int(w), int(h), inp["K"][i], inp["R"][i], inp["t"][i], inp["median_depth"][i], inp["wps"][i], inp["segs"][i])
        assert rc in (0, None), rc


def sample_points(segs, step=0.02):
    """points every `step` scene units along every 3D segment"""
    out = []
    for s in segs:
        a, b = s[:3], s[3:]
        n = max(2, int(np.linalg.norm(b - a) / step) + 1)
        out.append(a[None] + np.linspace(0, 1, n)[:, None] * (b - a)[None])
    return np.concatenate(out)


def chamfer(A, B):
    """median / 90th percentile distance from points A to the nearest point of B (scipy cKDTree)"""
    from scipy.spatial import cKDTree
    d, _ = cKDTree(B).query(A)
    return float(np.median(d)), float(np.quantile(d, 0.9))


def load_opt_pairs():
    """known-answer vectors of the reference's own Ceres bundling (tests/golden/make_opt_fixture.py): per cluster the line
    before and after LineOptimizer::optimize and its 2D residuals"""
    z = np.load(os.path.join(G, "line3dpp_ref_opt_pairs_v1.npz"))
    return z["before"], z["after"], z["res_ptr"], z["residuals"]


def optimizer_inputs(inp):
    """cams block (16 doubles per camera: R, C, fx, fy, px, py) in the reference's working frame, i.e. shifted by the
    per-axis median of the camera centres (Line3D::translate, line3D.cc:500-536); returns (cams, shift)"""
    V = inp["V"]
    cams = np.zeros((V, 16))
    for i in range(V):
        K, R, t = inp["K"][i], inp["R"][i], inp["t"][i]
        cams[i, :9] = R.reshape(9)
        cams[i, 9:12] = -R.T @ t
        cams[i, 12:16] = (K[0, 0], K[1, 1], K[0, 2], K[1, 2])
    C = cams[:, 9:12]
    shift = -np.array([np.sort(C[:, k])[V // 2] for k in range(3)])
    cams[:, 9:12] += shift
    return cams, shift


def dist_points_to_lines(P, L):
    """distance of point P[i] to the infinite line through L[i] = (a, b)"""
    c = L[:, :3]
    d = L[:, 3:] - L[:, :3]
    d = d / np.linalg.norm(d, axis=1, keepdims=True)
    w = P - c
    return np.linalg.norm(w - (w * d).sum(1, keepdims=True) * d, axis=1)


def line_gap(A, B):
    """max distance of the two end points of A[i] to the infinite line B[i]"""
    return np.maximum(dist_points_to_lines(A[:, :3], B), dist_points_to_lines(A[:, 3:], B))
