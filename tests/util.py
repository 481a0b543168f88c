"""Shared input preparation for the parity tests (numpy only; no arithmetic that is under test)."""
import numpy as np

from line3dpp_b200 import synth


def pair_inputs(scene, src, tgt):
    """Float inputs of one view pair exactly as matchingGPU hands them to match_lines_GPU (line3D.cc:1040-1064):
    float RtKinv / C of both views, float F.  No translation (kernel-level tests do not need it)."""
    RtKinv, C = synth.camera_blocks(scene)
    F = synth.fundamental(scene.K[src], scene.R[src], scene.t[src], scene.K[tgt], scene.R[tgt], scene.t[tgt])
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    return dict(ls=scene.segs[src], lt=scene.segs[tgt], F=f32(F).reshape(9), Rs=f32(RtKinv[src]).reshape(9),
                Rt=f32(RtKinv[tgt]).reshape(9), Cs=f32(C[src]), Ct=f32(C[tgt]))


def scene_descs(scene, k=None):
    """l3d_view_desc[] for a scene (double camera blocks from numpy; float copies made by capi.make_view_descs)."""
    from line3dpp_b200 import capi
    RtKinv, C = synth.camera_blocks(scene)
    V = scene.num_views
    k = np.zeros(V, np.float32) if k is None else k
    return capi.make_view_descs(scene.cam_ids, [scene.width] * V, [scene.height] * V, [len(s) for s in scene.segs],
                                RtKinv, C, C, k, np.zeros(V, np.float32))


def pair_F(scene, pairs):
    out = np.zeros((len(pairs), 9), np.float32)
    for i, (s, t) in enumerate(pairs):
        out[i] = synth.fundamental(scene.K[s], scene.R[s], scene.t[s], scene.K[t], scene.R[t], scene.t[t]).astype(np.float32).reshape(9)
    return out


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def rows_as_sets(counts, recs, fields=("tgt_seg", "overlap", "d_p1", "d_p2", "d_q1", "d_q2")):
    """per row: sorted list of tuples with float fields as raw bits"""
    out = []
    for r in range(len(counts)):
        row = []
        for i in range(counts[r]):
            e = recs[r, i]
            row.append(tuple(int(np.float32(e[f]).view(np.uint32)) if recs.dtype[f].kind == "f" else int(e[f]) for f in fields))
        out.append(sorted(row))
    return out


def two_view_scene(kind, n, seed):
    """two views of the same random 3D lines with an epipolar geometry the ring scenes do not have; kind = one of the named
    geometries or an explicit [(R, C), (R, C)] camera pair"""
    import dataclasses
    base = synth.make_scene(2, n, seed, "dense")
    rng = np.random.default_rng(seed)
    K = base.K[0]
    I = np.eye(3)
    rz = lambda a: np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1.0]])
    ry = lambda a: np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]])
    cams = kind if not isinstance(kind, str) else {
            "sideways": [(I, (-0.3, 0.0, -4.0)), (I, (0.3, 0.0, -4.0))],            # epipole at infinity (E.z == 0), horizontal epipolar lines
            "forward": [(I, (0.0, 0.0, -4.6)), (I, (0.04, -0.03, -3.7))],            # epipole inside the image
            "edge": [(I, (0.0, 0.0, -4.2)), (ry(0.05), (0.9, 0.1, -3.6))],           # epipole a little outside the image border
            "rolled": [(I, (-0.4, 0.1, -4.0)), (rz(1.45) @ ry(-0.08), (0.5, -0.2, -4.1))]}[kind]
    P1, P2 = base.lines3d[:, :3], base.lines3d[:, 3:]
    Rs, ts, segs = [], [], []
    for R, C in cams:
        C = np.array(C)
        t = -R @ C
        X1, X2 = (R @ P1.T).T + t, (R @ P2.T).T + t
        u1 = X1[:, :2] / X1[:, 2:3] * synth.FOCAL + K[:2, 2] + rng.normal(scale=0.5, size=(len(P1), 2))
        u2 = X2[:, :2] / X2[:, 2:3] * synth.FOCAL + K[:2, 2] + rng.normal(scale=0.5, size=(len(P1), 2))
        ok = (X1[:, 2] > 0.1) & (X2[:, 2] > 0.1) & (np.linalg.norm(u1 - u2, axis=1) >= synth.MIN_LEN_PX)
        for u in (u1, u2):
            ok &= (u[:, 0] >= 0) & (u[:, 0] <= synth.WIDTH - 1) & (u[:, 1] >= 0) & (u[:, 1] <= synth.HEIGHT - 1)
        segs.append(np.ascontiguousarray(np.concatenate([u1, u2], axis=1)[ok][:n].astype(np.float32)))
        Rs.append(R); ts.append(t)
    return dataclasses.replace(base, R=np.array(Rs), t=np.array(ts), segs=segs)
