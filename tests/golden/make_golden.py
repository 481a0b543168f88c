"""Generates tests/golden/ref_kernels_v1.npz by running the UNMODIFIED reference kernels (oracle/_ref, built from
/root/reference by oracle/Makefile with -fmad=false) on a B200:

    gpurun -- python tests/golden/make_golden.py gpurun_out/ref_kernels_v1.npz      # then copy into tests/golden/

The file pins, for small deterministic inputs: K_match_lines (dense overlaps + depths), match_lines_GPU (kNN lists),
K_score_matches (scores), SparseMatrix + replicator_dynamics_diffusion_GPU (diffused COO) and performClustering
(labels).  The CPU tests check oracle/l3d_oracle.cc against it; the GPU tests check the product against the same
reference live.  Inputs are regenerated from seeds by the tests (synth.make_scene is deterministic) and also stored.
"""
import sys
import os

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from line3dpp_b200 import synth  # noqa: E402
from oracle import pyoracle as po  # noqa: E402
from tests import util  # noqa: E402


def scoring_inputs(sc, src, tgts, ref, epi=0.25, knn=10):
    """flatten the kNN matches of view `src` towards `tgts` exactly like scoringGPU (line3D.cc:1311-1355)"""
    RtKinv, C = synth.camera_blocks(sc)
    per_seg = [[] for _ in range(len(sc.segs[src]))]
    for t in tgts:
        pi = util.pair_inputs(sc, src, t)
        counts, out, _, _ = po.match_lines(ref.ref_match_lines, pi["ls"], pi["lt"], pi["F"], pi["Rs"], pi["Rt"], pi["Cs"], pi["Ct"], src, t, epi, knn)
        for r in range(len(counts)):
            for i in range(counts[r]):
                per_seg[r].append((t, int(out[r, i]["tgt_seg"]), float(out[r, i]["d_p1"]), float(out[r, i]["d_p2"])))
    ranges = np.full((len(per_seg), 2), -1, np.int32)
    m4, reg = [], []
    off = 0
    k_t = 0.001
    for r, lst in enumerate(per_seg):
        lst.sort(key=lambda e: (e[0], e[1]))
        if lst:
            ranges[r] = (off, off + len(lst) - 1)
            off += len(lst)
        x1, y1, x2, y2 = sc.segs[src][r].astype(np.float64)
        for (t, ts, d1, d2) in lst:
            m4.append((float(r), float(t), d1, d2))
            r1 = RtKinv[src] @ np.array([x1, y1, 1.0]); r1 /= np.linalg.norm(r1)
            r2 = RtKinv[src] @ np.array([x2, y2, 1.0]); r2 /= np.linalg.norm(r2)
            P1, P2 = C[src] + r1 * d1, C[src] + r2 * d2
            reg.append((np.linalg.norm(P1 - C[t]) * k_t, np.linalg.norm(P2 - C[t]) * k_t))
    return (np.array(m4, np.float32), ranges, np.array(reg, np.float32), RtKinv[src].astype(np.float32).reshape(9), C[src].astype(np.float32))


def main(out_path):
    ref = po.ref_lib("nofma")
    assert ref is not None and ref.ref_device_count() > 0, "needs oracle/_ref and a GPU"
    sc = synth.make_scene(6, 160, 77, "dense")
    g = {"scene_args": np.array([6, 160, 77]), "segs": np.stack(sc.segs)}
    for (s, t) in [(0, 1), (2, 4)]:
        pi = util.pair_inputs(sc, s, t)
        dep, ov, _ = po.match_dense(ref.ref_match_dense, pi["ls"], pi["lt"], pi["F"], pi["Rs"], pi["Rt"], pi["Cs"], pi["Ct"], 0.25)
        counts, out, total, _ = po.match_lines(ref.ref_match_lines, pi["ls"], pi["lt"], pi["F"], pi["Rs"], pi["Rt"], pi["Cs"], pi["Ct"], s, t, 0.25, 10)
        tag = f"p{s}{t}"
        for k2 in ("F", "Rs", "Rt", "Cs", "Ct"):
            g[f"{tag}_{k2}"] = pi[k2]
        g[f"{tag}_dense_dep"], g[f"{tag}_dense_ov"] = dep, ov
        g[f"{tag}_knn_counts"], g[f"{tag}_knn"] = counts, out
    m4, ranges, reg, R, Cc = scoring_inputs(sc, 0, [1, 2, 3], ref)
    scores, _ = po.score_matches(ref.ref_score_matches, sc.segs[0], m4, ranges, reg, R, Cc, 200.0, 0.0011, 0.5)
    g.update(score_m4=m4, score_ranges=ranges, score_reg=reg, score_R=R, score_C=Cc, score_params=np.array([200.0, 0.0011, 0.5], np.float32),
             score_out=scores)
    rng = np.random.default_rng(5)
    n = 400
    a, b = rng.integers(0, n, 3000), rng.integers(0, n, 3000)
    keep = a != b
    a, b = a[keep], b[keep]
    _, idx = np.unique(np.minimum(a, b) * n + np.maximum(a, b), return_index=True)
    a, b = a[np.sort(idx)], b[np.sort(idx)]
    missing = np.setdiff1d(np.arange(n), np.concatenate([a, b]))
    a, b = np.concatenate([a, missing]), np.concatenate([b, (missing + 1) % n])
    w = rng.uniform(0.5, 1.0, len(a)).astype(np.float32)
    ei = np.stack([a, b], 1).reshape(-1).astype(np.int32)
    ej = np.stack([b, a], 1).reshape(-1).astype(np.int32)
    ew = np.repeat(w, 2)
    oi, oj, ow, _ = po.rdd(ref.ref_rdd, ei, ej, ew, n)
    lab = po.cluster(ref.ref_cluster, ei, ej, ew, n, 3.0)
    lab_rdd = po.cluster(ref.ref_cluster, oi, oj, ow, n, 3.0)
    g.update(rdd_n=np.array([n]), rdd_ei=ei, rdd_ej=ej, rdd_ew=ew, rdd_oi=oi, rdd_oj=oj, rdd_ow=ow, cluster_labels=lab, cluster_labels_rdd=lab_rdd)
    np.savez_compressed(out_path, **g)
    print("wrote", out_path, {k: v.shape for k, v in g.items()})


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/ref_kernels_v1.npz")
