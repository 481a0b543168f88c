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


def fixture_clusters():
    """the reference's own result file as known-answer vectors for the cluster -> 3D segment tail (findCollinearSegments,
    line3D.cc:2342-2452): per final line the infinite line through its 3D segments, its residuals (camera, per-camera segment
    index into `cam_segs`) and the expected 3D segments.  Returns (cam_segs, clusters)."""
    inp = load_inputs()
    segs3d, seg_line, res = load_fixture()
    cam_segs = [[] for _ in range(inp["V"])]
    seg_id = {}
    for r in res:
        key = (int(r[1]), int(r[2]))
        if key not in seg_id:
            seg_id[key] = len(cam_segs[key[0]])
            cam_segs[key[0]].append(r[3:7])
    cam_segs = [np.array(c, np.float32) if c else np.zeros((1, 4), np.float32) for c in cam_segs]
    clusters = []
    order = np.argsort(res[:, 0], kind="stable")
    bounds = np.searchsorted(res[order, 0], np.arange(int(res[:, 0].max()) + 2))
    for ln in sorted(set(seg_line.tolist())):
        S = segs3d[seg_line == ln]
        pts = np.concatenate([S[:, :3], S[:, 3:]])
        c = pts.mean(0)
        d = np.linalg.svd(pts - c)[2][0]
        t = (pts - c) @ d
        rr = res[order[bounds[ln]:bounds[ln + 1]]]
        clusters.append(dict(p1p2=np.concatenate([c + d * t.min(), c + d * t.max()]), cams=rr[:, 1].astype(np.uint32),
                             segs=np.array([seg_id[(int(a), int(b))] for a, b in rr[:, 1:3]], np.uint32), expect=S))
    return cam_segs, clusters


def check_fixture_segments(run, clusters):
    """run(cluster) -> (n,6) 3D segments; every segment of the reference output must be reproduced (text rounding ~5e-6)"""
    total = matched = extra = 0
    worst = 0.0
    for cl in clusters:
        mine = run(cl)
        for s in cl["expect"]:
            total += 1
            if len(mine) == 0:
                continue
            e = np.minimum(np.abs(mine - s).max(1), np.abs(mine - np.r_[s[3:], s[:3]]).max(1)).min()
            if e < 5e-5:
                matched += 1
                worst = max(worst, e)
        extra += max(0, len(mine) - len(cl["expect"]))
    return total, matched, extra, worst
