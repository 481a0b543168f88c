"""REF_CPU semantics on the GPU (SURVEY.md §8(b) dual-semantics switch; rows a3 / a7): the reference's CPU twins
matchingCPU (line3D.cc:900-1015) and scoringCPU (1208-1294) computed by the B200, checked against the CPU oracle's
restatement of them (oracle/l3d_oracle.cc: orc_match_lines_f64, scoring_cpu).  The matching part uses IEEE double
add/mul/div/sqrt only and is compared BIT FOR BIT; scoring goes through expf/acosf, whose CUDA and glibc versions differ by
ulps: scores are compared with a relative TOLERANCE of 1e-5, and decisions that sit on a threshold may flip for < 1 % of
the matches."""
import numpy as np
import pytest

from line3dpp_b200 import line3d, synth
from tests import util

pytestmark = pytest.mark.gpu


def _pair_inputs_f64(scene, s, t):
    RtKinv, C = synth.camera_blocks(scene)
    F = synth.fundamental(scene.K[s], scene.R[s], scene.t[s], scene.K[t], scene.R[t], scene.t[t])
    return dict(ls=scene.segs[s], lt=scene.segs[t], F=F.reshape(9), Rs=RtKinv[s].reshape(9), Rt=RtKinv[t].reshape(9), Cs=C[s], Ct=C[t])


@pytest.mark.parametrize("knn", [10, 3])
def test_f64_topk_bit_exact_vs_oracle(gpu_ctx, oracle, knn):
    scene = synth.make_scene(6, 900, 21, "ring2")
    pairs = synth.view_pairs(scene.neighbors)
    Fd = np.stack([synth.fundamental(scene.K[s], scene.R[s], scene.t[s], scene.K[t], scene.R[t], scene.t[t]).reshape(9) for s, t in pairs])
    gpu_ctx.set_views(util.scene_descs(scene), scene.segs)
    gpu_ctx.match_pairs_f64(pairs, Fd, 0.25, knn)
    assert gpu_ctx.L.l3d_match_semantics(gpu_ctx.h) == 1
    n_total = 0
    for pi, (s, t) in enumerate(pairs):
        inp = _pair_inputs_f64(scene, s, t)
        oc, om, ototal, _ = oracle.match_lines(oracle.lib().orc_match_lines_f64, inp["ls"], inp["lt"], inp["F"], inp["Rs"], inp["Rt"],
                                               inp["Cs"], inp["Ct"], scene.cam_ids[s], scene.cam_ids[t], 0.25, knn, f64=True)
        gc, gr = gpu_ctx.pair_matches(pi, len(inp["ls"]))
        assert np.array_equal(gc, oc), f"pair {pi}: per-row match counts differ"
        assert util.rows_as_sets(gc, gr) == util.rows_as_sets(oc, om), f"pair {pi}: kNN members / overlaps / depths differ"
        n_total += int(oc.sum())
    assert n_total > 2000


def test_f64_differs_from_f32_semantics(gpu_ctx):
    """the two semantics are different functions (SURVEY §8a divergences 1, 4): make sure the switch switches something"""
    scene = synth.make_scene(4, 1200, 22, "ring1")
    pairs = synth.view_pairs(scene.neighbors)
    Fd = np.stack([synth.fundamental(scene.K[s], scene.R[s], scene.t[s], scene.K[t], scene.R[t], scene.t[t]).reshape(9) for s, t in pairs])
    gpu_ctx.set_views(util.scene_descs(scene), scene.segs)
    gpu_ctx.match_pairs_f64(pairs, Fd, 0.25, 10)
    c64, r64 = gpu_ctx.pair_matches(0, len(scene.segs[pairs[0][0]]))
    gpu_ctx.match_pairs(pairs, Fd.astype(np.float32), 0.25, 10)
    assert gpu_ctx.L.l3d_match_semantics(gpu_ctx.h) == 0
    c32, r32 = gpu_ctx.pair_matches(0, len(scene.segs[pairs[0][0]]))
    same_members = sum(sorted(r64[r, :c64[r]]["tgt_seg"]) == sorted(r32[r, :c32[r]]["tgt_seg"]) for r in range(len(c64)))
    assert same_members > 0.97 * len(c64)                       # same geometry ...
    assert util.rows_as_sets(c64, r64) != util.rows_as_sets(c32, r32)   # ... different arithmetic


@pytest.fixture(scope="module")
def scene():
    return synth.make_scene(8, 500, 31, "ring3")


def test_refcpu_pipeline_vs_oracle(scene, oracle):
    L = line3d.Line3D(neighbors_by_worldpoints=False, use_gpu=False)
    L.add_scene(scene)
    L.match_images()
    P = oracle.OraclePipeline(False, False)          # use_gpu = False: matchingCPU + scoringCPU
    P.add_scene(scene)
    P.match_images()
    flips = total = 0
    for cam in scene.cam_ids:
        g, o = L.view_matches(cam, False), P.scored(cam)
        # same matches in the same (unsorted, append-order) list order, bit-identical geometry
        assert len(g) == len(o)
        for f in ("src_seg", "tgt_cam", "tgt_seg"):
            assert np.array_equal(g[f], o[f]), f"view {cam}: list order / membership differs in {f}"
        for f in ("overlap", "d_p1", "d_p2", "d_q1", "d_q2"):
            assert np.array_equal(util.bits(g[f]), util.bits(o[f])), f"view {cam}: {f} not bit-identical"
        bad = ~np.isclose(g["score3D"], o["score3D"], rtol=1e-5, atol=1e-6)     # TOLERANCE: libm vs libdevice expf/acosf
        flips += int(bad.sum()); total += len(g)
    assert total > 5000 and flips <= 0.01 * total, f"{flips} of {total} scores outside tolerance"
    # kept matches and estimates
    gb, gp = L.estimates()
    ob, op = P.estimates()
    gk = {(int(m["src_cam"]), int(m["src_seg"])): i for i, m in enumerate(gb)}
    ok_ = {(int(m["src_cam"]), int(m["src_seg"])): i for i, m in enumerate(ob)}
    common = set(gk) & set(ok_)
    assert len(common) >= 0.99 * max(len(gk), len(ok_)) and len(common) > 1000
    same_best = [k for k in common if gb[gk[k]]["tgt_cam"] == ob[ok_[k]]["tgt_cam"] and gb[gk[k]]["tgt_seg"] == ob[ok_[k]]["tgt_seg"]]
    assert len(same_best) >= 0.99 * len(common)
    a = np.array([gp[gk[k]] for k in same_best]); b = np.array([op[ok_[k]] for k in same_best])
    assert np.array_equal(a, b), "3D estimates of identical best matches must be bit-identical (double unprojection)"
    # reconstruction: no diffusion in REF_CPU (line3D.cc:1729); same lines up to the rare threshold flips
    L.reconstruct_3d_lines(3, True)
    assert P.reconstruct(3, True) == 0
    nl, no = L.stats()["lines3D"], P.num_lines()
    assert no > 100 and abs(nl - no) <= max(2, 0.02 * no)
    L.close()
