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
