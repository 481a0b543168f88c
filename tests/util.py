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
